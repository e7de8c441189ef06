# Evidence session: launch list of one bench step, ncu --set full of the attention kernel and of every HBM kernel (as CSV),
# CUDA-event timings of the hot launch classes, the default bench line.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2t_bench_default.log 2>&1
timeout 300 python tools/profile_attn.py --time > gpurun_out/r2t_attn_time.log 2>&1
timeout 300 python tools/profile_kernels.py --time > gpurun_out/r2t_tc_time.log 2>&1
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2t_hbm_time.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 2 -c 2 -o gpurun_out/r2t_prof_attn -f python tools/profile_attn.py > gpurun_out/r2t_ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_kernel|attn_tc_kernel" -c 8 -o /tmp/r2t_prof_tc -f python tools/profile_kernels.py > gpurun_out/r2t_ncu_tc.log 2>&1
ncu -i /tmp/r2t_prof_tc.ncu-rep --page raw --csv > gpurun_out/r2t_tc_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -o /tmp/r2t_prof_hbm -f python tools/profile_hbm.py > gpurun_out/r2t_ncu_hbm.log 2>&1
ncu -i /tmp/r2t_prof_hbm.ncu-rep --page raw --csv > gpurun_out/r2t_hbm_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --profile --steps 1 > gpurun_out/r2t_launches.log 2>&1
cat gpurun_out/r2t_attn_time.log gpurun_out/r2t_tc_time.log gpurun_out/r2t_hbm_time.log; tail -1 gpurun_out/r2t_bench_default.log | cut -c1-200; du -sh gpurun_out
