set -x
TAG=${TAG:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/kernel_bench.py > gpurun_out/${TAG}_kb.jsonl 2> gpurun_out/${TAG}_kb.err; tail -2 gpurun_out/${TAG}_kb.err
timeout 200 python tools/chain_profile.py > gpurun_out/${TAG}_chain.jsonl 2> gpurun_out/${TAG}_chain.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --kernel-table > gpurun_out/${TAG}_bench_akt1.json 2> gpurun_out/${TAG}_bench_akt1.err; tail -2 gpurun_out/${TAG}_bench_akt1.err
python - <<'P'
import json, os
tag = os.environ.get("TAG", "ab")
for l in open(f"gpurun_out/{tag}_kb.jsonl"):
    try:
        d = json.loads(l); print({k: d[k] for k in list(d)[:6]})
    except Exception: pass
d = json.loads(open(f"gpurun_out/{tag}_bench_akt1.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d["gpu_launches"], d["peak_mem_gb"], d["losses"])
for k, v in list(d["kernel_table"].items())[:12]: print(k, v)
P
