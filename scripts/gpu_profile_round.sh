#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench, ncu launch list and ncu --set full captures for the round's profiles.
# usage: bash scripts/gpu_profile_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/${TAG}_status.log
python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/${TAG}_status.log 2>&1
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_status.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_10steps.json 2>> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
# every launch of the bench command (pipelined steps with the lane-per-stream coder, then the serial accounting pass) with
# its device time -- cold-cache and serialised under ncu: compare shares
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --inflight 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# full capture: one batch through the path, per coder layout (the second pass of each; the first warms up)
# (15 kernels per pass; skip the first warm-up pass; the report stays on the box -- only its raw-page CSV comes back)
ncu --set full --clock-control none -k regex:"tc_|rans_|nchw" -s 15 -c 45 -o /tmp/${TAG}_full -f \
    python scripts/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>> $OUT/${TAG}_profile_step.log
python scripts/diag_trace.py 8 48 > $OUT/${TAG}_trace_streams.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv
cat $OUT/${TAG}_status.log
