set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "ffn or conv2d" > gpurun_out/p9_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -4 gpurun_out/p9_pytest.log
timeout 300 python tools/profile_kernels.py --time > gpurun_out/p9_time.log 2>&1; cat gpurun_out/p9_time.log
if [ $rc -eq 0 ]; then
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p9_bench.log 2>&1
tail -1 gpurun_out/p9_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline_conv']['frac'], d['roofline']['frac'], d['sections_ms_per_step'], d['epe_vs_reference']['mean'])"
fi
