set -x
TAG=${TAG:-n3}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_native_block_gpu.py -q 2>&1 | tail -40 > gpurun_out/${TAG}_native_tests.log; tail -12 gpurun_out/${TAG}_native_tests.log
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_native_block_gpu.py 2>&1 | tail -30 > gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_notarget_b512.json 2> gpurun_out/${TAG}_bench_nt.err; tail -2 gpurun_out/${TAG}_bench_nt.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_akt1.json 2> gpurun_out/${TAG}_bench_akt1.err; tail -2 gpurun_out/${TAG}_bench_akt1.err
python - <<'P'
import json, glob, os
tag = os.environ.get("TAG", "n3")
for f in sorted(glob.glob(f"gpurun_out/{tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d["gpu_launches"], d["peak_mem_gb"], d["losses"],
              d.get("roofline", {}).get("kernel"), d.get("roofline", {}).get("launches"), d.get("roofline", {}).get("avg_launch_ms"))
    except Exception as e:
        print(f, "ERR", e)
P
