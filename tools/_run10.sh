set -x
mkdir -p gpurun_out
for m in thread warp thread warp; do
UM_PAIR_ARRIVE=$m timeout 300 python tools/profile_kernels.py --time 2>&1 | grep -E "gru|ffn" | tr '\n' '|'; echo " <- $m"
done
UM_PAIR_ARRIVE=thread timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p10_bench_thread.log 2>&1
UM_PAIR_ARRIVE=warp timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p10_bench_warp.log 2>&1
for f in thread warp; do tail -1 gpurun_out/p10_bench_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$f', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline_conv']['frac'], d['sections_ms_per_step'])"; done
