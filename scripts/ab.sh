# A/B of build variants: usage  AB_MODELS="deepfm xdeepfm" bash scripts/ab.sh libA.so libB.so ...   (per-kernel ms of every phase)
mkdir -p gpurun_out
for m in ${AB_MODELS:-deepfm}; do
for v in "$@"; do
  B200REC_LIB=$PWD/recommendation-models_b200/$v timeout 400 python bench.py --steps ${AB_STEPS:-30} --warmup 5 --no-cpu --model $m > gpurun_out/ab_${m}_$v.json 2> gpurun_out/ab_${m}_$v.err || tail -5 gpurun_out/ab_${m}_$v.err
  python - "$v" "$m" <<'PY'
import json, sys
v, m = sys.argv[1], sys.argv[2]
for l in open(f'gpurun_out/ab_{m}_{v}.json'):
    if l.startswith('{'):
        d = json.loads(l); print(m, v, d['value'], d['ms_per_step'])
        for k in d['kernels']: print('   ', k['phase'], k['ms_per_step'], {a: round(b, 4) for a, b in k['kernels'].items()})
PY
done; done 2>&1 | tee gpurun_out/${AB_TAG:-ab}.txt
