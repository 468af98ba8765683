# round-2 final evidence run on one B200: tests, the driver's bench command, the reference arm, ncu launch lists,
# ncu --set full of the top kernels (summarised on the box: the reports themselves exceed gpurun's 64 MiB)
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02zz_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zz_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02zz_ref.json 2> gpurun_out/r02zz_ref.err; echo "ref rc=$?"
for m in deepfm xdeepfm; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02zz_launches_$m.csv python bench.py --model $m --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/r02zz_ncu_l_$m.log 2>&1; echo "ncu launches $m rc=$?"
done
timeout 900 ncu --set full --clock-control none -k regex:"gemm_ws_kernel|cin_dz_kernel|segsum_kernel|fm_fwd_kernel|emb_grad_kernel" -c 24 -o /tmp/r02zz_full_xdeepfm -f python bench.py --model xdeepfm --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/r02zz_ncu_fx.log 2>&1; echo "ncu full x rc=$?"
python profiles/summarize.py full /tmp/r02zz_full_xdeepfm.ncu-rep > gpurun_out/r02zz_full_xdeepfm.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_ws_kernel|segsum_kernel|fm_fwd_kernel|emb_grad_kernel|head_bwd|pack_|rs_|seg_" -c 44 -o /tmp/r02zz_full_deepfm -f python bench.py --model deepfm --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/r02zz_ncu_fd.log 2>&1; echo "ncu full d rc=$?"
python profiles/summarize.py full /tmp/r02zz_full_deepfm.ncu-rep > gpurun_out/r02zz_full_deepfm.txt 2>&1
du -sh gpurun_out; tail -3 gpurun_out/r02zz_pytest.log
