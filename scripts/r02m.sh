mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r02m_pytest_sharded.log 2>&1; echo "pytest rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_bench_2.json 2> gpurun_out/r02m_bench_2.err; echo "bench2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r02m_ref_2.json 2> gpurun_out/r02m_ref_2.err; echo "ref2 rc=$?"
tail -3 gpurun_out/r02m_pytest_sharded.log; tail -c 600 gpurun_out/r02m_bench_2.err
