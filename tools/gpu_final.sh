# Final measurement set of a round (bench lines, reference arm, kernel micro-bench, chain phases, encoder sweep).  TAG=r02e bash tools/gpu_final.sh
set -x
TAG=${TAG:-r02e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/${TAG}_gpu_tests.log; tail -2 gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 5 --warmup 3 --kernel-table > gpurun_out/${TAG}_bench_n1_akt1_b2048.json 2> gpurun_out/${TAG}_bench_akt1.err
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --kernel-table --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_notarget_b512.json 2>/dev/null
python bench.py --precision bf16x3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_akt1_b2048_bf16x3.json 2> gpurun_out/${TAG}_bench_bf16x3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2>/dev/null
python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.jsonl 2>/dev/null
python tools/chain_profile.py > gpurun_out/${TAG}_chain_phases.jsonl 2>/dev/null
timeout 150 python tools/encoder_sweep.py --batches 256,2048,16384 --atoms 9,45 --heads 8 > gpurun_out/${TAG}_encoder_sweep.jsonl 2>/dev/null
python - <<'P'
import json, glob, os
tag = os.environ.get("TAG", "r02e")
for f in sorted(glob.glob(f"gpurun_out/{tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d.get("gpu_launches"), d.get("peak_mem_gb"), d.get("losses"),
              {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("roofline", {}).items() if k in ("kernel", "frac", "tensor_frac", "hbm_frac_implemented", "share_of_step", "avg_launch_ms", "launches")},
              d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
P
