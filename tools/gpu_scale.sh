# 8-GPU box: multi-slab parity at world 8 (one case per halo path), then the weak-scaling bench at N = 8 and N = 1.
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "5-even-8-1 or 5-odd-8-1 or 5-even-8-nccl" > gpurun_out/pytest_gpu8.log 2>&1; echo rc=$?
grep -E "passed|failed|halo_p2p" gpurun_out/pytest_gpu8.log | sort | uniq -c | tail -6
for n in 8 1; do
  timeout 400 python bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  python - <<PY
import json
j=json.load(open("gpurun_out/scale_n$n.json"))
print($n, "value", round(j["value"],1), "ms/step", round(j["ms_per_step"],5), "bulk frac", round(j["roofline"]["frac"],4), "e2e", round(j["e2e"]["value"],1), "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"], j["config"]["partition"][:50])
PY
done
