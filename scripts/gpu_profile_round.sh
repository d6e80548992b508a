#!/bin/bash
# Runs on the GPU box (under gpurun): tests, smoke, bench lines of all five configs, the reference arm, the ncu launch list of the bench
# command and ncu --set full captures of the transforms and the coder.  usage: bash scripts/gpu_profile_round.sh <tag>
set -u
TAG=${1:-r4z}
OUT=gpurun_out
mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee $OUT/${TAG}_status.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/${TAG}_status.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_status.log
timeout 400 python bench.py > $OUT/${TAG}_bench_config2.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_status.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
for c in 1 3 4 5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > $OUT/${TAG}_bench_config$c.json 2>> $OUT/${TAG}_bench.err
done
# every launch of the bench command (pipelined steps, then the serial accounting pass) with its device time -- cold-cache and
# serialised under ncu: compare shares
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --inflight 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# full capture of one pass of the transforms (the third; the first two warm up): 9 kernels
timeout 400 ncu --set full --clock-control none -k regex:"tc_|ga_|nchw" -s 18 -c 9 -o /tmp/${TAG}_tr -f python scripts/prof_transforms.py > $OUT/${TAG}_prof.log 2>&1
ncu -i /tmp/${TAG}_tr.ncu-rep --page raw --csv > $OUT/${TAG}_tr_raw.csv 2>> $OUT/${TAG}_prof.log
# the coder: both layouts (profile_step.py: two passes per layout, the second of each is captured)
timeout 400 ncu --set full --clock-control none -k regex:"rans_(en|de)code" -c 8 -o /tmp/${TAG}_coder -f python scripts/profile_step.py >> $OUT/${TAG}_prof.log 2>&1
ncu -i /tmp/${TAG}_coder.ncu-rep --page raw --csv > $OUT/${TAG}_coder_raw.csv 2>> $OUT/${TAG}_prof.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv
cat $OUT/${TAG}_status.log; tail -3 $OUT/${TAG}_pytest.log
