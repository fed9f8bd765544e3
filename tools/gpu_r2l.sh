python -m pytest tests/test_gpu_tb.py -m gpu -x -q 2>&1 | tail -2
python tools/tb_sweep.py slab 2 2 120; python tools/tb_sweep.py slab 2 2 120; python tools/tb_sweep.py c4 2 2 40
