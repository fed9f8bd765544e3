python -m pytest tests/test_gpu_tb.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log | cut -c1-300
for pf in 0 1 2 4; do for B in 256 128; do
  LBM_B200_TB_PF=$pf LBM_B200_TB_B=$B LBM_B200_TB_XC=64 python tools/tb_sweep.py slab 2 2 120 | sed "s/^{/{\"pf\": $pf, /"
done; done 2>&1 | grep -v "^$" | tee gpurun_out/r2e_sweep.jsonl
LBM_B200_TB_B=128 python tools/tb_sweep.py slab 2 2 120; python tools/tb_sweep.py c4 2 2 40
CMD="python tools/tb_sweep.py slab 2 2 20"
LBM_B200_TB_B=128 LBM_B200_TB_XC=64 $CMD > gpurun_out/r2e_plain.log 2>&1 &&
LBM_B200_TB_B=128 LBM_B200_TB_XC=64 ncu --set full --clock-control none --import-source on -k regex:k_tb -s 4 -c 1 -f -o gpurun_out/r2e_tb2 $CMD > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"
