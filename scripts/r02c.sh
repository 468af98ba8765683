mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 6 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02c_pytest.log
