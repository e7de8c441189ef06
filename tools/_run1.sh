set -x
mkdir -p gpurun_out
UM_CONV_PAIR=1 timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "conv2d" > gpurun_out/p1_pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/p1_pytest_conv.log
UM_CONV_PAIR=0 timeout 300 python tools/profile_kernels.py --time > gpurun_out/p1_time_pair0.log 2>&1
UM_CONV_PAIR=1 timeout 300 python tools/profile_kernels.py --time > gpurun_out/p1_time_pair1.log 2>&1
paste gpurun_out/p1_time_pair0.log gpurun_out/p1_time_pair1.log
