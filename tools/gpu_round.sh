# End-of-round check on one B200: parity suite, smoke, the judged bench lines, the other workloads.
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final_slab.json 2> gpurun_out/final_slab.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "ref rc=$?"
for cfg in "c1" "c3" "slab"; do for aa in "" "--aa"; do
  python bench.py --workload $cfg $aa --no-cpu-baseline > "gpurun_out/final_${cfg}${aa}.json" 2> /dev/null; done; done
for aa in "" "--aa"; do python bench.py --workload c4 $aa --no-cpu-baseline --no-e2e --steps 300 > "gpurun_out/final_c4${aa}.json" 2> /dev/null; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j.get("roofline") or {}
    e = j.get("e2e") or {}
    print(f.split("final_")[1], "value", round(j["value"], 1), "ms/step", j.get("ms_per_step") and round(j["ms_per_step"], 5),
          "frac", r.get("frac") and round(r["frac"], 4), "whole", j.get("roofline_whole_step_frac") and round(j["roofline_whole_step_frac"], 4),
          "e2e", e.get("value") and round(e["value"], 1), "cpu", (j.get("cpu_baseline") or {}).get("value"))
PY
