#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), probes.  Usage: gpurun -- 'bash tools/gpu_call.sh <tag> [steps...]'
TAG=${1:-x}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt; cat /sys/fs/cgroup/cpu.max >> $OUT/${TAG}_smi.txt 2>&1
for step in "$@"; do
  case $step in
    tests)  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.err ;;
    ref)    timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$?" ;;
    probe)  timeout 120 ./tools/l2_reduce_probe > $OUT/${TAG}_l2probe.txt 2>&1; echo "probe rc=$?"; cat $OUT/${TAG}_l2probe.txt ;;
    *) echo "unknown step $step" ;;
  esac
done
