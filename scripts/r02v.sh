# A/B of environment switches on the xDeepFM bench: usage  bash scripts/r02v.sh "VAR=1" "VAR=0 OTHER=2" ...
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model ${AB_MODEL:-xdeepfm} > gpurun_out/abenv_$i.json 2> gpurun_out/abenv_$i.err
  python - "$v" $i <<'PY'
import json, sys
v, i = sys.argv[1], sys.argv[2]
for l in open(f'gpurun_out/abenv_{i}.json'):
    if l.startswith('{'):
        d = json.loads(l); print(v, d['value'], d['ms_per_step'])
        for k in d['kernels'][:2]: print('   ', k['phase'], k['ms_per_step'], {a: b for a, b in k['kernels'].items() if b > 0.05})
PY
done 2>&1 | tee gpurun_out/r02v_ab.txt
