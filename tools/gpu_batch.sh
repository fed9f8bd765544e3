python -m pytest tests/test_gpu_extensions.py -m gpu -x -q 2>&1 | tail -8
for w in c3 c4 c1; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
