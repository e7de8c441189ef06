# One GPU-box session: parity first, then the bench line, timings and profiles (outputs under gpurun_out/r2a_*).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
# second-generation attention kernel first; fall back to the first generation for the rest of the session if it fails
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "fullsize_vs_oracle or test_window_attention" -p no:cacheprovider > gpurun_out/r2a_attn_v2.log 2>&1
if ! tail -3 gpurun_out/r2a_attn_v2.log | grep -q " passed" || tail -3 gpurun_out/r2a_attn_v2.log | grep -q "failed"; then export UM_ATTN_V1=1; echo "FALLBACK to v1" >> gpurun_out/r2a_attn_v2.log; fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_stages_gpu.py -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$? UM_ATTN_V1=$UM_ATTN_V1" >> gpurun_out/r2a_pytest.log
timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_stages.log 2>&1; echo "stages rc=$?" >> gpurun_out/r2a_stages.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2a_bench.log
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2a_hbm_time.log 2>&1
timeout 300 python tools/profile_kernels.py --time > gpurun_out/r2a_tc_time.log 2>&1
UM_ATTN_V1=1 timeout 300 python tools/profile_kernels.py --time > gpurun_out/r2a_tc_time_v1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -o gpurun_out/r2a_prof_hbm -f python tools/profile_hbm.py > gpurun_out/r2a_ncu_hbm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --profile --steps 1 > gpurun_out/r2a_launches.log 2>&1
tail -5 gpurun_out/r2a_attn_v2.log; tail -5 gpurun_out/r2a_pytest.log; tail -30 gpurun_out/r2a_stages.log; tail -3 gpurun_out/r2a_bench.log; cat gpurun_out/r2a_tc_time.log gpurun_out/r2a_tc_time_v1.log
