# round-2 final multi-GPU evidence: usage  bash scripts/r02zz_multi.sh N   (sharded checks on N ranks, then the driver's bench command at N)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/sharded_check.py > gpurun_out/r02zz_sharded_check_world$N.txt 2>&1; echo "sharded_check rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02zz_bench_${N}gpu_100m.json 2> gpurun_out/r02zz_bench_${N}gpu.err; echo "bench rc=$?"
tail -4 gpurun_out/r02zz_sharded_check_world$N.txt
python - <<PY
import json
for l in open("gpurun_out/r02zz_bench_${N}gpu_100m.json"):
    if l.startswith("{"):
        d = json.loads(l); x = d.get("xdeepfm") or {}
        print("N=$N deepfm", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| xdeepfm", x.get("value"), x.get("ms_per_step"))
        print(d.get("kernels_rank0_ms_per_step"))
PY
