# compute-sanitizer memcheck + racecheck + synccheck over the kernel unit tests (small row counts only: the instrumented
# kernels run 10-100x slower).  Summaries -> gpurun_out/${TAG}_sanitizer.txt   (TAG=r02c bash tools/sanitize.sh)
mkdir -p gpurun_out
OUT=gpurun_out/${TAG:-r02}_sanitizer.txt
: > $OUT
SEL='not 19021 and not 37893 and not 56837 and not 47365 and not 40-45 and not 300-9 and not 40000 and not 100000 and not 56829 and not depth8 and not 1000003 and not 300-45 and not 150-17 and not 2048 and not 20000'
N=0
run() {   # tool, file, extra -k
  N=$((N+1))
  echo "== compute-sanitizer --tool $1  python -m pytest $2 -m gpu -k '$SEL $3'" >> $OUT
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 5 \
      python -m pytest $2 -q -m gpu -k "$SEL $3" > gpurun_out/san_${N}_$1.log 2>&1
  echo "   exit code $? (124 = the per-file time limit cut the run; what ran is summarised below)" >> $OUT
  grep -E "passed|failed|^FAILED|ERROR SUMMARY|RACECHECK SUMMARY|========= (Error|Race|Barrier)" gpurun_out/san_${N}_$1.log | sort | uniq -c | tail -8 >> $OUT
}
run memcheck tests/test_kernels_gpu.py ""
run memcheck tests/test_attn_chain_gpu.py ""
run memcheck tests/test_glue_gpu.py ""
run memcheck tests/test_data_gpu.py "and not full_batch and not golden_and_properties"
run racecheck tests/test_kernels_gpu.py "and (fused_mlp or bf16_storage or accumulating)"
run racecheck tests/test_attn_chain_gpu.py "and not l2_prefetch"     # (R = 37965 under racecheck exceeds the kernels' bounded mbarrier wait)
run synccheck tests/test_kernels_gpu.py "and (fused_mlp or bf16_storage or accumulating)"
run synccheck tests/test_attn_chain_gpu.py ""
run racecheck tests/test_data_gpu.py "and not full_batch and not golden_and_properties"
cat $OUT
