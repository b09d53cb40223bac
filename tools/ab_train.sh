#!/bin/bash
# usage: tools/ab_train.sh <lib suffix> ...   (A/B of the training step: forward in training mode + staged backward)
for v in "$@"; do
  GMPI_LIB_PATH=$PWD/ml_gmpi_b200/libgmpi_mpi_render_$v.so timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d['train_step']
print('$v', 'fwd', round(d['value'], 1), 'fps; train', round(t['value'], 1), 'fps', round(t['ms_per_step'], 3), 'ms/step')"
done
