# compute-sanitizer memcheck + racecheck + synccheck over the kernel unit tests (small row counts only: the instrumented
# kernels run 10-100x slower).  Summaries -> gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer.txt
: > $OUT
SEL='not 19021 and not 37893 and not 40-45 and not depth8'
for TOOL in memcheck racecheck synccheck; do
  for FILE in tests/test_kernels_gpu.py tests/test_attn_chain_gpu.py tests/test_glue_gpu.py; do
    echo "== compute-sanitizer --tool $TOOL  python -m pytest $FILE -m gpu -k '$SEL'" >> $OUT
    timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 5 \
        python -m pytest $FILE -q -m gpu -x -k "$SEL" > gpurun_out/san_$TOOL.log 2>&1
    echo "   exit code $?" >> $OUT
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" gpurun_out/san_$TOOL.log | tail -6 >> $OUT
  done
done
cat $OUT
