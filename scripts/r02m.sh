# 2-GPU box: sharded checks (world 1 and 2), bench --gpus 2 on the 100 M-row table (DeepFM only for the A/B)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r02m_pytest_sharded.log 2>&1; echo "pytest rc=$?"
for v in 1 0; do
B200REC_FUSED_PUSH=$v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$v bench.py --gpus 2 --steps 30 --warmup 5 --model deepfm > gpurun_out/r02m_bench_2_push$v.json 2> gpurun_out/r02m_bench_2_push$v.err; echo "bench2 push=$v rc=$?"
done
tail -3 gpurun_out/r02m_pytest_sharded.log; tail -c 300 gpurun_out/r02m_bench_2_push1.err
