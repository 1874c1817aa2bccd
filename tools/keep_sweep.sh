# A/B of the activation policy (block.keep_intermediates): safety margin sweep + off.   bash tools/keep_sweep.sh
for h in ${HEADROOMS:-40 70 100}; do
  DRUGGEN_B200_KEEP_HEADROOM_GB=$h python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headroom $h', round(d['value'],1), round(d['ms_per_step'],1), d['peak_mem_gb'], d['reserved_mem_gb'], d['allocator'], d['gpu_launches'])"
done
DRUGGEN_B200_KEEP=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('keep off', round(d['value'],1), round(d['ms_per_step'],1), d['peak_mem_gb'], d['reserved_mem_gb'], d['allocator'], d['gpu_launches'])"
