set -x
TAG=${TAG:-n2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 200 python tools/encoder_sweep.py --batches 256,2048 --atoms 9,45 --heads 8 > gpurun_out/${TAG}_encoder_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err
DRUGGEN_B200_GRAPH=0 timeout 200 python tools/encoder_sweep.py --batches 256,2048 --atoms 9,45 --heads 8 > gpurun_out/${TAG}_encoder_sweep_nograph.jsonl 2>/dev/null
DRUGGEN_B200_NATIVE_BLOCK=0 timeout 200 python tools/encoder_sweep.py --batches 256,2048 --atoms 9,45 --heads 8 > gpurun_out/${TAG}_encoder_sweep_py.jsonl 2>/dev/null
python bench.py --batch 512 --workload NoTarget --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_notarget_b512.json 2> gpurun_out/${TAG}_bench_nt.err; tail -2 gpurun_out/${TAG}_bench_nt.err
cut -c1-200 gpurun_out/${TAG}_encoder_sweep*.jsonl
