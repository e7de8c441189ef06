set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider > gpurun_out/r2m_ops.log 2>&1
timeout 120 python tools/profile_attn.py --time > gpurun_out/r2m_attn_time.log 2>&1
timeout 300 python tools/profile_kernels.py --time > gpurun_out/r2m_tc_time.log 2>&1
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2m_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2m_bench.log 2>&1
tail -3 gpurun_out/r2m_ops.log; cat gpurun_out/r2m_attn_time.log gpurun_out/r2m_tc_time.log; tail -3 gpurun_out/r2m_e2e.log; tail -2 gpurun_out/r2m_bench.log | cut -c1-300
