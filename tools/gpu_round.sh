set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "fullsize_vs_oracle or test_window_attention" -p no:cacheprovider > gpurun_out/r2f_attn.log 2>&1
timeout 300 python tools/profile_attn.py --time > gpurun_out/r2f_attn_time.log 2>&1
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2f_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2f_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 2 -c 2 -o gpurun_out/r2f_prof_attn -f python tools/profile_attn.py > gpurun_out/r2f_ncu_attn.log 2>&1
tail -15 gpurun_out/r2f_attn.log; cat gpurun_out/r2f_attn_time.log; tail -4 gpurun_out/r2f_e2e.log; tail -2 gpurun_out/r2f_bench.log | cut -c1-400
