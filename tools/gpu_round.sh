set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv
timeout 900 python -m pytest tests/test_attn_chain_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/t_chain.log; tail -5 gpurun_out/t_chain.log
timeout 1500 python -m pytest tests -q -m gpu --ignore tests/test_attn_chain_gpu.py 2>&1 | tail -40 > gpurun_out/t_all.log; tail -5 gpurun_out/t_all.log
timeout 600 python tools/kernel_bench.py --prefetch-ab > gpurun_out/kb.jsonl 2> gpurun_out/kb.err; tail -3 gpurun_out/kb.err
timeout 600 python bench.py --steps 3 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
DRUGGEN_B200_L2_PREFETCH=0 timeout 600 python bench.py --steps 3 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/bench_pf0.json 2> gpurun_out/bench_pf0.err
DRUGGEN_B200_ATTN_CHAIN=0 timeout 600 python bench.py --steps 3 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/bench_chain0.json 2> gpurun_out/bench_chain0.err
python - <<'P'
import json
for f in ("bench_a","bench_pf0","bench_chain0"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["peak_mem_gb"], d["losses"])
    except Exception as e: print(f, "ERR", e)
P
