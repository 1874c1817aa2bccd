set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/t_all.log; tail -4 gpurun_out/t_all.log
timeout 300 python tools/chain_profile.py > gpurun_out/chain_prof.jsonl 2> gpurun_out/chain_prof.err
timeout 600 python tools/kernel_bench.py ${KB_ARGS:-} > gpurun_out/kb.jsonl 2> gpurun_out/kb.err; tail -3 gpurun_out/kb.err
timeout 600 python bench.py --steps 3 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<'P'
import json
for f in ("bench_a",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["peak_mem_gb"], d["losses"])
    except Exception as e: print(f, "ERR", e)
P
