set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "local_corr" -p no:cacheprovider > gpurun_out/r2q_ops.log 2>&1
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2q_hbm_time.log 2>&1
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2q_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2q_bench.log 2>&1
UM_ATTN_DEBUG_BUILD=1 timeout 600 python tools/attn_timeline.py --s1 > gpurun_out/r2q_timeline_s1.log 2>&1
tail -3 gpurun_out/r2q_ops.log; head -6 gpurun_out/r2q_hbm_time.log; tail -3 gpurun_out/r2q_e2e.log; tail -1 gpurun_out/r2q_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sections_ms_per_step'], d['epe_vs_reference']['pass'])"
sed -n 1,24p gpurun_out/r2q_timeline_s1.log
