# round 2, call C: leaner temporally blocked kernel -- parity, sweep, ncu
python -m pytest tests/test_gpu_tb.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log | cut -c1-300
for B in 256 128; do for xc in 0 32 64 128; do
  if [ $xc = 0 ]; then LBM_B200_TB_B=$B python tools/tb_sweep.py slab 2 2 120; else LBM_B200_TB_B=$B LBM_B200_TB_XC=$xc python tools/tb_sweep.py slab 2 2 120; fi
done; done 2>&1 | grep -v "^$" | tee gpurun_out/r2c_sweep.jsonl
python tools/tb_sweep.py slab 2 1 120; python tools/tb_sweep.py slab 1 1 120; python tools/tb_sweep.py slab 2 3 120
CMD="python tools/tb_sweep.py slab 2 2 20"
$CMD > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tb -s 4 -c 1 -f -o gpurun_out/r2c_tb2 $CMD > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"
