set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_config4.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload config2 > gpurun_out/r2n_bench_config2.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload config3 > gpurun_out/r2n_bench_config3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload config5 > gpurun_out/r2n_bench_config5.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --graph --no-cpu-baseline > gpurun_out/r2n_bench_graph.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_bench_reference.log 2>&1
tail -4 gpurun_out/r2n_pytest.log; for f in config4 config2 config3 config5 graph reference; do tail -1 gpurun_out/r2n_bench_$f.log | cut -c1-260; done
