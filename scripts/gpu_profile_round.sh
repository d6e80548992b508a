#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench, ncu launch list and ncu --set full captures for the round's profiles.
# usage: bash scripts/gpu_profile_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/${TAG}_status.log
python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/${TAG}_status.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_status.log
python bench.py --steps 10 --warmup 3 --inflight 1 --no-e2e --no-cpu-baseline > $OUT/${TAG}_bench_serial.json 2>> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
# every launch of one serial step with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --inflight 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# full capture of one step's own kernels (skip the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:"tc_|rans_|patchify|nchw" -s 64 -c 18 -o $OUT/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --inflight 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv
cat $OUT/${TAG}_status.log
