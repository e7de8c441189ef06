set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "ffn or gelu" > gpurun_out/p8_pytest_ffn.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -5 gpurun_out/p8_pytest_ffn.log
UM_FFN_WIDE=0 timeout 300 python tools/profile_kernels.py --time > gpurun_out/p8_time_wide0.log 2>&1
UM_FFN_WIDE=1 timeout 300 python tools/profile_kernels.py --time > gpurun_out/p8_time_wide1.log 2>&1
paste gpurun_out/p8_time_wide0.log gpurun_out/p8_time_wide1.log
