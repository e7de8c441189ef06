set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2e_bench_default.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/r2e_bench_default.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline_conv']['frac'], d['sections_ms_per_step'], d['epe_vs_reference']['mean'], d['epe_vs_reference']['pass'], d['gpu_launches'])"
