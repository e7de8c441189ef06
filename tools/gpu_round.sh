set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "fullsize_vs_oracle or test_window_attention" -p no:cacheprovider > gpurun_out/r2h_attn.log 2>&1
for f in 0 16 1 8; do echo "== UM_ATTN_DBG=$f" >> gpurun_out/r2h_attn_time.log; UM_ATTN_DBG=$f timeout 120 python tools/profile_attn.py --time >> gpurun_out/r2h_attn_time.log 2>&1; done
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2h_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2h_bench.log 2>&1
tail -15 gpurun_out/r2h_attn.log; cat gpurun_out/r2h_attn_time.log; tail -4 gpurun_out/r2h_e2e.log; tail -2 gpurun_out/r2h_bench.log | cut -c1-400
