set -x
mkdir -p gpurun_out
timeout 300 python tools/profile_attn.py --time > gpurun_out/r2b_attn_time.log 2>&1
UM_ATTN_V1=1 timeout 300 python tools/profile_attn.py --time > gpurun_out/r2b_attn_time_v1.log 2>&1
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2b_hbm_time.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -o gpurun_out/r2b_prof_attn -f python tools/profile_attn.py > gpurun_out/r2b_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -o gpurun_out/r2b_prof_hbm -f python tools/profile_hbm.py > gpurun_out/r2b_ncu_hbm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --profile --steps 1 > gpurun_out/r2b_launches.log 2>&1
cat gpurun_out/r2b_attn_time.log gpurun_out/r2b_attn_time_v1.log gpurun_out/r2b_hbm_time.log; ls -la gpurun_out
