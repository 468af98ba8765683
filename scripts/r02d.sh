mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
ncu --set full --clock-control none --import-source on -k regex:"segsum_kernel|fm_fwd_kernel" -s 6 -c 4 -o gpurun_out/r02d_segsum -f python bench.py --model deepfm --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02d_pytest.log
