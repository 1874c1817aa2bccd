# Round measurements: tests, bench lines (throughput mode, tensor-core parity mode, NoTarget), reference arm, kernel micro-bench,
# chain phase profile, encoder sweep, ncu --set full of the hot kernels.   TAG=r02 bash tools/gpu_measure.sh
set -x
TAG=${TAG:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log; tail -2 gpurun_out/${TAG}_tests.log
python bench.py --steps 5 --warmup 3 --kernel-table > gpurun_out/${TAG}_bench_n1_akt1_b2048.json 2> gpurun_out/${TAG}_bench_akt1.err
python bench.py --precision bf16x3 --steps 3 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_akt1_b2048_bf16x3.json 2> gpurun_out/${TAG}_bench_bf16x3.err
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_notarget_b512.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2>/dev/null
python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.jsonl 2>/dev/null
python tools/chain_profile.py > gpurun_out/${TAG}_chain_phases.jsonl 2>/dev/null
timeout 240 python tools/encoder_sweep.py --batches 256,2048,16384 > gpurun_out/${TAG}_encoder_sweep.jsonl 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_chain|rows_gemm_tc|gemm_tn_tc|attn_scores|attn_fwd_warp|add_ln_bwd_kernel|bwd_bwd' -s 21 -c 21 -o gpurun_out/${TAG}_prof python tools/profile_one.py > gpurun_out/${TAG}_ncu_prof.log 2>&1
ls -la gpurun_out/ | grep " ${TAG}_"
