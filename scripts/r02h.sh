mkdir -p gpurun_out
for cfg in "4 8" "8 8" "8 13" "12 13" "16 20"; do
  set -- $cfg
  rm -f gpurun_out/parity_report.jsonl
  B200REC_KC=$1 B200REC_KC_SHORT=$2 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02h_bench_kc$1_s$2.json 2> gpurun_out/r02h_bench_kc$1_s$2.err
  B200REC_KC=$1 B200REC_KC_SHORT=$2 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_models.py -m gpu -q -k "benchmark_batch or baseline_sized" > gpurun_out/r02h_pytest_kc$1_s$2.log 2>&1
  cp gpurun_out/parity_report.jsonl gpurun_out/r02h_parity_kc$1_s$2.jsonl
  tail -1 gpurun_out/r02h_pytest_kc$1_s$2.log
done
