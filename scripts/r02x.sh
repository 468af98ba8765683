# Round 2, campaign x: (1) what paces the DeepFM MLP GEMMs -- timing-only build variants (wrong results by
# construction): B loader copying half the bytes, producers without their shared-memory stores, both;
# (2) the random-row-gather ceiling of this GPU (scripts/ubench/gather_rate.cu).
mkdir -p gpurun_out
for v in libb200rec.so libb200rec_halfb.so libb200rec_noastore.so libb200rec_halfb_noastore.so; do
  B200REC_LIB=$PWD/recommendation-models_b200/$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --model deepfm > gpurun_out/x_$v.json 2> gpurun_out/x_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
for l in open(f'gpurun_out/x_{v}.json'):
    if l.startswith('{'):
        d = json.loads(l); print(v, d['value'], d['ms_per_step'])
        for k in d['kernels']: print('   ', k['phase'], k['ms_per_step'], k['kernels'])
PY
done 2>&1 | tee gpurun_out/r02x_ab.txt
timeout 120 ./scripts/ubench/gather_rate 2>&1 | tee gpurun_out/r02x_gather_rate.txt
