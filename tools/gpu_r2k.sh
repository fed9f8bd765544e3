for u in 1 0; do for wl in slab c4; do
  st=120; [ $wl = c4 ] && st=40
  LBM_B200_TB_UNROLL=$u python tools/tb_sweep.py $wl 2 2 $st | sed "s/^{/{\"unroll\": $u, /"
done; done
LBM_B200_TB_XC=48 python tools/tb_sweep.py slab 2 2 120
LBM_B200_TB_XC=96 python tools/tb_sweep.py slab 2 2 120
