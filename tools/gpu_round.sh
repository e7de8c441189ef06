set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "local_corr" -p no:cacheprovider > gpurun_out/r2r_ops.log 2>&1
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2r_hbm_time.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/r2r_bench.log 2>&1
tail -3 gpurun_out/r2r_ops.log; head -7 gpurun_out/r2r_hbm_time.log; tail -1 gpurun_out/r2r_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sections_ms_per_step'])"
