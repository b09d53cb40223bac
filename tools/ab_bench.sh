#!/bin/bash
# usage: tools/ab_bench.sh <lib suffix> ...   (A/B kernel builds: ml_gmpi_b200/libgmpi_mpi_render_<suffix>.so)
for v in "$@"; do
  GMPI_LIB_PATH=$PWD/ml_gmpi_b200/libgmpi_mpi_render_$v.so timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-train-step 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['value'], 1), 'fps', round(d['roofline']['frac'], 4), 'of roofline', round(d['roofline']['kernel_ms'], 4), 'ms')"
done
