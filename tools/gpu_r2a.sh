# round 2, call A: temporal-blocking parity on one B200, a first bench line, the launch-shape sweep
python -m pytest tests/test_gpu_tb.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2a_bench.json
python tools/tb_sweep.py slab > gpurun_out/r2a_sweep.jsonl 2> gpurun_out/r2a_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r2a_sweep.jsonl
