set -x
mkdir -p gpurun_out
for f in 0 1 2 4 6 7 3 5; do echo "== UM_ATTN_DBG=$f" >> gpurun_out/r2e_attn_dbg.log; UM_ATTN_DBG=$f timeout 120 python tools/profile_attn.py --time >> gpurun_out/r2e_attn_dbg.log 2>&1; done
timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "local_corr_softmax" -p no:cacheprovider > gpurun_out/r2e_local.log 2>&1
timeout 300 python tools/profile_hbm.py --time > gpurun_out/r2e_hbm_time.log 2>&1
cat gpurun_out/r2e_attn_dbg.log; tail -5 gpurun_out/r2e_local.log; head -5 gpurun_out/r2e_hbm_time.log
