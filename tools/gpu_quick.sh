mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_attn_chain_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/chain_profile.py > gpurun_out/chain_prof.jsonl 2> gpurun_out/chain_prof.err
timeout 600 python tools/kernel_bench.py > gpurun_out/kb.jsonl 2> gpurun_out/kb.err; tail -3 gpurun_out/kb.err
