set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" 
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
