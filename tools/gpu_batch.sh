python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$?; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
for pdl in 1 0; do for cfg in "c1" "c1 --aa" "c3" "slab"; do LBM_B200_PDL=$pdl python bench.py --workload $cfg --no-cpu-baseline --no-e2e --steps 2000 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('pdl=$pdl $cfg', round(j['value'],1), 'ms/step', round(j['ms_per_step'],5), 'bulk ms', round(j['roofline']['avg_launch_ms'],5), 'whole', round(j['roofline_whole_step_frac'],4))"; done; done
