# round 2, call H (1 GPU): the whole GPU suite after the engine changes
python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2h_pytest_gpu.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload c4 --no-cpu-baseline --no-e2e --steps 300 > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err; echo "c4 rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2h_bench_c4.json')); print(j['value'], j['roofline']['frac'], j.get('physics_check'))"
python bench.py --workload c3 --no-cpu-baseline --steps 1000 > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/r2h_bench_c3.json
