set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o_bench_2gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --graph > gpurun_out/r2o_bench_2gpu_graph.log 2>&1
tail -2 gpurun_out/r2o_bench_2gpu.log | cut -c1-3000; tail -1 gpurun_out/r2o_bench_2gpu_graph.log | cut -c1-400
