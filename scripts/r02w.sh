mkdir -p gpurun_out
run() { # label, env..., -- bench args
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model xdeepfm $EXTRA > gpurun_out/w_$label.json 2> gpurun_out/w_$label.err
  python - $label <<'PY'
import json, sys
for l in open(f'gpurun_out/w_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d = json.loads(l); print(sys.argv[1], d['value'], d['ms_per_step'])
        for k in d['kernels'][:2]: print('   ', k['phase'], k['ms_per_step'], {a: b for a, b in k['kernels'].items() if b > 0.5})
PY
}
EXTRA="--gemm-mode 2" run tf32x1 A=1
EXTRA="" run noastore B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_noast.so
EXTRA="--gemm-mode 2" run noastore_tf32x1 B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_noast.so
EXTRA="--gemm-mode 2" run halfb_tf32x1 B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_halfb.so
