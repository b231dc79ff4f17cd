#!/bin/bash
# decode throughput per fusion mask (see engine.cu fuse_mask): usage tools/mask_sweep.sh <workload> <steps> <masks...>
wl=$1; steps=$2; shift 2
python bench.py --workload $wl --steps $steps --warmup 4 --no-fused --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl unfused', round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches'])"
for m in "$@"; do
UZU_FUSE_MASK=$m python bench.py --workload $wl --steps $steps --warmup 4 --fused --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl mask $m', round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches'])"
done
