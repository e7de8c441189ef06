set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "ffn_tc or bn96" > gpurun_out/p7_pytest_ffn.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -25 gpurun_out/p7_pytest_ffn.log
if [ $rc -eq 0 ]; then
timeout 900 python -m pytest tests/test_module_gpu.py tests/test_stages_gpu.py -x -q -m gpu > gpurun_out/p7_pytest_mod.log 2>&1; echo "pytest mod rc=$?"
tail -3 gpurun_out/p7_pytest_mod.log
for f in 1 0; do
UM_FUSED_FFN=$f timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p7_bench_ffn$f.log 2>&1
tail -1 gpurun_out/p7_bench_ffn$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('fused $f', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline_conv']['frac'], d['sections_ms_per_step'], d['epe_vs_reference']['mean'])"
done
fi
