# Round 2, campaign y: source-level ncu capture (stall samples per SASS line) of the first two MLP forward GEMMs
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_ws_kernel -c 2 -o gpurun_out/r02y_gemm_src -f python bench.py --model deepfm --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/r02y_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
