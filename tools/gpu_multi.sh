# Multi-GPU evidence (under gpurun --gpus N): bash tools/gpu_multi.sh N
#   N = 2: the whole multi-slab matrix (temporally blocked passes of depth 1/2/3, NCCL fallback, periodic x, observers with
#          skewed ranks, the silent-neighbour time-out, AA over slabs), the 2-slab C++ driver, bench at N = 2
#   N = 4, 8: a subset of the matrix at world N and N/2, bench at N (and the in-place AA variant)
N=${1:-2}
if [ "$N" = 2 ]; then
  python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_driver.py -m gpu -q > gpurun_out/multi_pytest_world2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/multi_pytest_world2.log | cut -c1-400
else
  K="5-even-$N-1-2-2 or 5-odd-$N-1-2-2 or 5-even-$N-1-2-3 or 5-even-$N-nccl-2-2 or 5-even-37-$N or 5-odd-37-$N"
  python -m pytest tests/test_gpu_multi.py -m gpu -q -k "$K" > gpurun_out/multi_pytest_world$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/multi_pytest_world$N.log | cut -c1-400
  python bench.py --gpus $N --aa --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/multi_bench_n${N}_aa.json 2> /dev/null; echo "bench aa rc=$?"
fi
nvidia-smi topo -m > gpurun_out/multi_topo_n$N.txt 2>&1
python bench.py --gpus $N --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err; echo "bench rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/multi_bench_n$N.json')); e=j['e2e']
print('N=$N', round(j['value']), 'MLUPS', round(j['ms_per_step'],4), 'ms/step, e2e', round(e['value']), e['rank0_ms'], j['multi_gpu_parity'])"
