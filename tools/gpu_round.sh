# Evidence session (one GPU): GPU test suite, bench lines of every single-GPU workload, CUDA-event timings of the hot launch
# classes, ncu metrics of the CTA-pair kernels, launch list of one bench step.  Results land in gpurun_out/r2d_*.
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2d_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_default.log 2>&1
for c in config2 config3 config5; do timeout 600 python bench.py --workload $c --steps 5 --warmup 3 > gpurun_out/r2d_bench_$c.log 2>&1; done
timeout 300 python tools/profile_pair.py --time > gpurun_out/r2d_pair_time.log 2>&1
timeout 300 python tools/profile_kernels.py --time > gpurun_out/r2d_tc_time.log 2>&1
timeout 300 python tools/profile_attn.py --time > gpurun_out/r2d_attn_time.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"conv_tc_kernel|ffn_tc_kernel" -s 4 -c 4 --csv --page raw --log-file gpurun_out/r2d_ncu_pair_raw.csv python tools/profile_pair.py > gpurun_out/r2d_ncu_pair.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --profile --steps 1 > gpurun_out/r2d_launches.log 2>&1
cat gpurun_out/r2d_pair_time.log gpurun_out/r2d_tc_time.log gpurun_out/r2d_attn_time.log
for f in default config2 config3 config5; do tail -1 gpurun_out/r2d_bench_$f.log | cut -c1-300; done
du -sh gpurun_out
