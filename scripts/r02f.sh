mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
./scripts/ubench/mma_rate > gpurun_out/r02f_mma_rate.txt 2>&1; echo "ubench rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 6 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02f_pytest.log
