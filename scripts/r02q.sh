# 8-GPU box: sharded checks on 8 ranks, bench --gpus 8 and --gpus 4 (100 M-row table, both models)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r02q_pytest_sharded.log 2>&1; echo "pytest rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02q_bench_8.json 2> gpurun_out/r02q_bench_8.err; echo "bench8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02q_bench_4.json 2> gpurun_out/r02q_bench_4.err; echo "bench4 rc=$?"
tail -3 gpurun_out/r02q_pytest_sharded.log; tail -c 400 gpurun_out/r02q_bench_8.err
