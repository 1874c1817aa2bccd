# Round measurements: bench lines, reference arm, ncu launch list, ncu --set full of the hot kernels, kernel micro-bench.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --kernel-table > gpurun_out/m_bench_akt1.json 2> gpurun_out/m_bench_akt1.err
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/m_bench_notarget.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/m_bench_ref.json 2>/dev/null
python tools/kernel_bench.py > gpurun_out/m_kb.jsonl 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 6000 --csv --log-file gpurun_out/m_launches.csv python bench.py --batch 512 --workload NoTarget --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/m_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_chain|rows_gemm_tc|gemm_tn_tc|attn_scores|attn_fwd_warp|add_ln_bwd_kernel|bwd_bwd' -s 18 -c 18 -o gpurun_out/m_prof python tools/profile_one.py > gpurun_out/m_ncu_prof.log 2>&1
ls -la gpurun_out/ | tail -12
