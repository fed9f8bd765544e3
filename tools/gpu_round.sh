# End-of-round check on one B200: the whole GPU suite, smoke, the judged bench lines, the other workloads, ncu evidence.
# Usage (under gpurun): [WITH_REFERENCE=1] [WITH_AA=1] bash tools/gpu_round.sh
python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_20.json 2> gpurun_out/final_bench_20.err; echo "bench20 rc=$?"
[ -n "$WITH_REFERENCE" ] && { python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"; }
for wl in c3 c4 c1; do python bench.py --workload $wl --no-cpu-baseline $( [ $wl = c4 ] && echo "--no-e2e --steps 300" ) > gpurun_out/final_bench_$wl.json 2>/dev/null; echo "$wl rc=$?"; done
python bench.py --variant 1 --no-cpu-baseline --no-e2e --steps 400 > gpurun_out/final_bench_variant1.json 2>/dev/null; echo "v1 rc=$?"
python bench.py --depth 2 --no-cpu-baseline --no-e2e --steps 400 > gpurun_out/final_bench_depth2.json 2>/dev/null; echo "depth2 rc=$?"
[ -n "$WITH_AA" ] && { python bench.py --aa --no-cpu-baseline --no-e2e --steps 400 > gpurun_out/final_bench_aa.json 2>/dev/null; echo "aa rc=$?"; }
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try: j = json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    r = j.get("roofline") or {}; e = j.get("e2e") or {}
    print(f.split("final_bench_")[1], "value", round(j["value"], 1), "frac", r.get("frac") and round(r["frac"], 3), "x144", r.get("frac_at_144B_per_update") and round(r["frac_at_144B_per_update"], 3), "share", r.get("kernel_share_of_step") and round(r["kernel_share_of_step"], 3),
          "e2e", e.get("value") and round(e["value"], 1), "cpu", (j.get("cpu_baseline") or {}).get("value"), "parity", (j.get("parity_check") or {}).get("bit_identical"), (j.get("parity_check") or {}).get("sha_matches_oracle"))
PY
bash tools/ncu_round.sh r02 > gpurun_out/final_ncu.log 2>&1; tail -2 gpurun_out/final_ncu.log
