set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "fullsize_vs_oracle or test_window_attention" -p no:cacheprovider > gpurun_out/r2s_attn.log 2>&1
timeout 120 python tools/profile_attn.py --time > gpurun_out/r2s_attn_time.log 2>&1
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2s_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2s_bench.log 2>&1
tail -3 gpurun_out/r2s_attn.log; cat gpurun_out/r2s_attn_time.log; tail -3 gpurun_out/r2s_e2e.log; tail -1 gpurun_out/r2s_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_class'], d['epe_vs_reference']['pass'])"
