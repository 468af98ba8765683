# A/B of build variants on the xDeepFM bench: usage  bash scripts/r02u.sh libA.so libB.so ...
mkdir -p gpurun_out
for v in "$@"; do
  B200REC_LIB=$PWD/recommendation-models_b200/$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model ${AB_MODEL:-xdeepfm} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
for l in open(f'gpurun_out/ab_{v}.json'):
    if l.startswith('{'):
        d = json.loads(l); print(v, d['value'], d['ms_per_step'])
        for k in d['kernels']: print('   ', k['phase'], k['ms_per_step'], {a: b for a, b in k['kernels'].items() if b > 0.05})
PY
done 2>&1 | tee gpurun_out/r02u_ab.txt
