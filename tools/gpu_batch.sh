python -m pytest tests/test_gpu_aa.py -m gpu -x -q > gpurun_out/pytest_aa.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_aa.log | cut -c1-300
for w in slab c4; do python bench.py --workload $w --aa --no-cpu-baseline > gpurun_out/bench_aa_$w.json 2> gpurun_out/bench_aa_$w.err; python - <<PY
import json
j=json.load(open("gpurun_out/bench_aa_$w.json"))
print("$w AA value", round(j["value"],1), "ms/step", round(j["ms_per_step"],5), "bulk frac", round(j["roofline"]["frac"],4), "e2e", round(j["e2e"]["value"],1), j["stable"])
PY
tail -2 gpurun_out/bench_aa_$w.err; done
