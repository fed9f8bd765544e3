# 8-GPU box: multi-slab parity tests at world 2/4/8, then the weak-scaling bench at N = 1, 2, 4, 8.
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_gpu8.log
for n in 1 2 4 8; do
  python bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  python - <<PY
import json
j=json.load(open("gpurun_out/scale_n$n.json"))
print($n, "value", round(j["value"],1), "ms/step", round(j["ms_per_step"],5), "bulk frac", round(j["roofline"]["frac"],4), "e2e", round(j["e2e"]["value"],1), "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
PY
done
