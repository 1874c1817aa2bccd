# Round measurements: tests, bench lines, reference arm, kernel micro-bench, chain phase profile, encoder sweep, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/m_tests.log; tail -2 gpurun_out/m_tests.log
python bench.py --steps 5 --warmup 3 --kernel-table > gpurun_out/m_bench_akt1.json 2> gpurun_out/m_bench_akt1.err
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/m_bench_notarget.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/m_bench_ref.json 2>/dev/null
python tools/kernel_bench.py > gpurun_out/m_kb.jsonl 2>/dev/null
python tools/chain_profile.py > gpurun_out/m_chain_prof.jsonl 2>/dev/null
timeout 240 python tools/encoder_sweep.py --batches 256,2048,16384 > gpurun_out/m_sweep.jsonl 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_chain|rows_gemm_tc|gemm_tn_tc|attn_scores|attn_fwd_warp|add_ln_bwd_kernel|bwd_bwd' -s 18 -c 18 -o gpurun_out/m_prof python tools/profile_one.py > gpurun_out/m_ncu_prof.log 2>&1
ls -la gpurun_out/ | grep " m_"
