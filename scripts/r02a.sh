mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02a_launches_xdeepfm.csv python bench.py --model xdeepfm --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/r02a_ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"cin_dz_kernel|gemm_ws_kernel" -c 20 -o gpurun_out/r02a_full_xdeepfm -f python bench.py --model xdeepfm --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/r02a_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/r02a_pytest.log
