#!/bin/bash
# Retry a gpurun call until it is neither "transient" (no slot) nor refused; at most $2 attempts, 100 s apart.
CMD="$1"; N=${2:-12}; shift; shift
for i in $(seq 1 $N); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" -- "$CMD" 2>&1)
  echo "$OUT" | tail -40
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry] attempt $i transient; sleeping"; sleep 100; continue; fi
  break
done
