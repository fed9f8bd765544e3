python -m pytest tests/test_gpu_tb.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log | cut -c1-300
for pf in 1 2; do for B in 256 128; do for xc in 32 64 96; do
  LBM_B200_TB_PF=$pf LBM_B200_TB_B=$B LBM_B200_TB_XC=$xc python tools/tb_sweep.py slab 2 2 120 | sed "s/^{/{\"pf\": $pf, /"
done; done; done 2>&1 | grep -v "^$" | tee gpurun_out/r2f_sweep.jsonl
python tools/tb_sweep.py slab 2 2 120; python tools/tb_sweep.py c4 2 2 40; python tools/tb_sweep.py c3 2 2 200; python tools/tb_sweep.py c1 2 2 1000; python tools/tb_sweep.py c1 1 1 1000
