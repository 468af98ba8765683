mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 6 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_trace.so timeout 300 python scripts/trace_tc.py > gpurun_out/r02i_timeline.txt 2>&1
tail -4 gpurun_out/r02i_pytest.log
