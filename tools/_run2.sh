set -x
mkdir -p gpurun_out
UM_CONV_PAIR=1 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/p2_pytest_gpu.log
UM_CONV_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p2_bench_pair1.log 2>&1
UM_CONV_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p2_bench_pair0.log 2>&1
UM_CONV_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p2_bench_pair1b.log 2>&1
for f in gpurun_out/p2_bench_pair1.log gpurun_out/p2_bench_pair0.log gpurun_out/p2_bench_pair1b.log; do tail -1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline_conv']['frac'], d['roofline']['frac'], d['sections_ms_per_step'], d['epe_vs_reference']['mean'])"; done
