# source-level ncu capture of a CIN forward GEMM (layer 2, H=200: 273 stages)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_ws_kernel --launch-skip 4 -c 1 -o gpurun_out/r02y_cin_src -f python bench.py --model xdeepfm --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/r02y2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
