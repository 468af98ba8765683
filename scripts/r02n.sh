mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r02n_pytest.log
