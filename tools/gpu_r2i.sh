# round 2, call I (2 GPUs): AA over slabs, 2-slab C++ driver with the shared-image async VTK, written f_next
python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "in_place_aa" > gpurun_out/r2i_pytest_aa_slabs.log 2>&1; echo "aa slabs rc=$?"; tail -12 gpurun_out/r2i_pytest_aa_slabs.log | cut -c1-600
python -m pytest tests/test_gpu_cpp_driver.py tests/test_gpu_aa.py -m gpu -q -x > gpurun_out/r2i_pytest_driver.log 2>&1; echo "driver rc=$?"; tail -12 gpurun_out/r2i_pytest_driver.log | cut -c1-600
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "written_f_next or golden" > gpurun_out/r2i_pytest_fnext.log 2>&1; echo "fnext rc=$?"; tail -8 gpurun_out/r2i_pytest_fnext.log | cut -c1-600
