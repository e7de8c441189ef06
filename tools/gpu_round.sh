set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -k "conv2d" -p no:cacheprovider > gpurun_out/r2p_ops.log 2>&1
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_module_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2p_e2e.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-gpu > gpurun_out/r2p_bench.log 2>&1
tail -3 gpurun_out/r2p_ops.log; tail -3 gpurun_out/r2p_e2e.log; tail -1 gpurun_out/r2p_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['sections_ms_per_step'], d['epe_vs_reference'])"
