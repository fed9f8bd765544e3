# round 2, call G (2 GPUs): the multi-slab matrix with the temporally blocked passes, bench at N=2, NVLink counters
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2g_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/r2g_pytest_multi.log | cut -c1-400
nvidia-smi nvlink -gt d > gpurun_out/r2g_nvlink_before.txt 2>&1
python bench.py --gpus 2 --steps 400 --warmup 10 --no-cpu-baseline > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-2500 gpurun_out/r2g_bench_n2.json
nvidia-smi nvlink -gt d > gpurun_out/r2g_nvlink_after.txt 2>&1
python bench.py --gpus 1 --steps 400 --warmup 10 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-3000 gpurun_out/r2g_bench_n1.json
