# round 2, call J (8 GPUs): multi-slab parity at world 8 and 4 (temporally blocked passes, NCCL fallback, AA), bench at N = 8 and 4
K8="5-even-8-1-2-2 or 5-odd-8-1-2-2 or 5-even-8-1-2-3 or 5-even-8-nccl-2-2 or 5-even-37-8 or 5-odd-37-8"
K4="5-even-4-1-2-2 or 5-odd-4-1-2-3 or 5-even-36-4"
python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "$K8 or $K4" > gpurun_out/r2j_pytest_world8.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest_world8.log | cut -c1-500
python bench.py --gpus 8 --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "bench n8 rc=$?"
python bench.py --gpus 4 --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err; echo "bench n4 rc=$?"
python bench.py --gpus 8 --aa --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_n8_aa.json 2> gpurun_out/r2j_bench_n8_aa.err; echo "bench n8 aa rc=$?"
python - <<'PY'
import json
for f in ("r2j_bench_n8", "r2j_bench_n4", "r2j_bench_n8_aa"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "unreadable", e); continue
    e2e = j.get("e2e") or {}
    print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 4), "e2e", e2e.get("value") and round(e2e["value"]), e2e.get("rank0_ms"),
          "parity", j.get("multi_gpu_parity"), "clocks", j.get("clocks", {}).get("sm_mhz"), j.get("clocks", {}).get("reasons"))
PY
tail -3 gpurun_out/r2j_bench_n8.err
