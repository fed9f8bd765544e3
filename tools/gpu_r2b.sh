# round 2, call B: ncu --set full of the temporally blocked kernel (depth 2, default shape), one GPU
CMD="python tools/tb_sweep.py slab 2 2 20"
$CMD > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tb -s 4 -c 2 -f -o gpurun_out/r2b_tb2 $CMD > gpurun_out/r2b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
