mkdir -p gpurun_out
B200REC_CIN_DW_T=1 TRACE_CIN_BWD=1 TRACE_CIN_BATCH=8192 B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_trace.so timeout 600 python scripts/trace_tc.py 2>&1 | sed -n '/cin dW/,$p' > gpurun_out/r02t_trace.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "cin or xdeepfm or encoder or step" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model xdeepfm > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02t_pytest.log; tail -3 gpurun_out/r02t_trace.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02t_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'])
        for k in d['kernels']: print(k['phase'], k['ms_per_step'], {a:b for a,b in k['kernels'].items() if b>0.1})
PY
