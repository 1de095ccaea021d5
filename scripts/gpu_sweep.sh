#!/bin/bash
# warp-count sweep of K1 on cfg2 / cfg3 (BPLX_NWARPS overrides the plan's choice)
for W in 10 16 20 24; do
  for wl in cfg2 cfg3; do
    BPLX_NWARPS=$W python bench.py --workload $wl --steps 50 --warmup 5 --no-extras --cpu-seconds 0.1 --cpu-chains 8 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('W=$W $wl us/step %.2f value %.4g smem %d' % (1e3*j['ms_per_step'], j['value'], j['config']['plan']['smem_bytes']))"
  done
done
