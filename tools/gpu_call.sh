#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), probes.  Usage: gpurun -- 'bash tools/gpu_call.sh <tag> [steps...]'
TAG=${1:-x}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt; cat /sys/fs/cgroup/cpu.max >> $OUT/${TAG}_smi.txt 2>&1
for step in "$@"; do
  case $step in
    tests)  timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.err ;;
    ref)    timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$?" ;;
    probe)  timeout 120 ./tools/l2_reduce_probe > $OUT/${TAG}_l2probe.txt 2>&1; echo "probe rc=$?"; cat $OUT/${TAG}_l2probe.txt ;;
    bwdtests) timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or grad or non_projective" > $OUT/${TAG}_pytest_bwd.log 2>&1; echo "bwd pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_bwd.log ;;
    trainbench) timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-on-gpu > $OUT/${TAG}_trainbench.json 2> $OUT/${TAG}_trainbench.err; echo "trainbench rc=$?"; python -c "import json;d=json.load(open('$OUT/${TAG}_trainbench.json'));print(d['value'],d['roofline']['frac'],d['train_step'],d['configs']['C5_train_512'])" ;;
    ncubwd) timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpi_bwd -c 1 -o $OUT/${TAG}_prof_bwd python tools/run_one.py bwd > $OUT/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?" ;;
    ncufwd) timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpi_fwd_staged -c 1 -o $OUT/${TAG}_prof_fwd python tools/run_one.py fwd > $OUT/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?" ;;
    sanitize) timeout 900 compute-sanitizer --tool racecheck python tools/run_one.py small > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/${TAG}_racecheck.log; timeout 900 compute-sanitizer --tool memcheck python tools/run_one.py small > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/${TAG}_memcheck.log ;;
    fwdab)  timeout 300 python tools/fwd_ab.py r02a - minimal scalarepi 2>&1 | tee $OUT/${TAG}_fwdab.txt ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-reference-on-gpu > $OUT/${TAG}_launches_bench.log 2>&1; echo "launches rc=$?" ;;
    multi)  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $OUT/${TAG}_pytest_multi.log 2>&1; echo "multi pytest rc=$?"; tail -4 $OUT/${TAG}_pytest_multi.log ;;
    benchn) NG=$(nvidia-smi -L | wc -l); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$NG.json 2> $OUT/${TAG}_bench_n$NG.err; echo "bench n=$NG rc=$?"; tail -c 800 $OUT/${TAG}_bench_n$NG.err; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $NG --steps 2 --warmup 1 --ref-budget-s 40 > $OUT/${TAG}_ref_n$NG.json 2> $OUT/${TAG}_ref_n$NG.err; echo "ref n=$NG rc=$?" ;;
    ncuextra) for w in fwdfact c4 c4flat; do timeout 400 ncu --set full --clock-control none -k regex:mpi_fwd_staged -c 1 -o $OUT/${TAG}_prof_$w python tools/run_one.py $w > $OUT/${TAG}_ncu_$w.log 2>&1; echo "ncu $w rc=$?"; done ;;
    abtrain) for rep in 1 2; do for v in $ABV; do GMPI_LIB_PATH=$PWD/ml_gmpi_b200/libgmpi_mpi_render_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reference-on-gpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'fwd', round(d['roofline']['kernel_ms'], 4), 'ms; train', round(d['train_step']['ms_per_step'], 3), 'ms', round(d['train_step']['roofline_frac'], 4), 'C5', round(d['configs']['C5_train_512']['ms_per_step'], 3), 'N1', round(d['configs']['N1_factored_fwd']['ms_per_step'], 4))" | tee -a $OUT/${TAG}_abtrain.txt; done; done ;;
    vtests) for v in ${VTV:-$ABV}; do GMPI_LIB_PATH=$PWD/ml_gmpi_b200/libgmpi_mpi_render_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -x -q -k "backward or grad or non_projective or factored or bwd" > $OUT/${TAG}_pytest_$v.log 2>&1; echo "$v pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_$v.log; done ;;
    zprobe) timeout 300 python tools/zero_overlap_probe.py > $OUT/${TAG}_zprobe.txt 2>&1; echo "zprobe rc=$?"; tail -2 $OUT/${TAG}_zprobe.txt ;;
    zab) for v in $ABV; do for fr in $FRACS; do GMPI_LIB_PATH=$PWD/ml_gmpi_b200/libgmpi_mpi_render_$v.so GMPI_ZERO_FRAC=$fr timeout 200 python tools/zero_ab.py 2>&1 | tail -1 | tee -a $OUT/${TAG}_zab.txt; done; done ;;
    *) echo "unknown step $step" ;;
  esac
done
