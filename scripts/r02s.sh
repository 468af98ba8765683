mkdir -p gpurun_out
for t in 1 0; do
  echo "=== B200REC_CIN_DW_T=$t" 
  B200REC_CIN_DW_T=$t TRACE_CIN_BWD=1 TRACE_CIN_BATCH=8192 B200REC_LIB=$PWD/recommendation-models_b200/libb200rec_trace.so timeout 600 python scripts/trace_tc.py 2>&1 | sed -n '/cin dW/,$p'
done > gpurun_out/r02s_trace.txt 2>&1
tail -5 gpurun_out/r02s_trace.txt
