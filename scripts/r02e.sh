mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 6 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"segsum_kernel" -s 2 -c 2 -o gpurun_out/r02e_segsum -f python bench.py --model deepfm --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02e_pytest.log
