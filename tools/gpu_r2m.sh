python -m pytest tests/test_gpu_tb.py -m gpu -x -q 2>&1 | tail -2
for sk in 1 0; do LBM_B200_TB_SKEW=$sk python tools/tb_sweep.py slab 2 2 120 | sed "s/^{/{\"skew\": $sk, /"; done
for pf in 0 2 3; do LBM_B200_TB_PF=$pf python tools/tb_sweep.py slab 2 2 120 | sed "s/^{/{\"skew\": 1, \"pf\": $pf, /"; done
LBM_B200_TB_B=256 python tools/tb_sweep.py slab 2 2 120
python tools/tb_sweep.py c4 2 2 40
python tools/tb_sweep.py slab 2 3 120
