#!/bin/bash
# Run on the GPU box (gpurun -- bash profiles/run_profiles.sh <tag>): launch lists + one full-set capture
# per dominant kernel.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/bench_under_ncu_$TAG.log 2>&1
# (2) full-set captures
ncu --set full --clock-control none --import-source on -k regex:allpairs_kernel -s 1 -c 1 -f -o $OUT/prof_gravity_$TAG \
    python profiles/prof_driver.py gravity > $OUT/prof_gravity_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_force_kernel -s 1 -c 1 -f -o $OUT/prof_lj_$TAG \
    python profiles/prof_driver.py lj > $OUT/prof_lj_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_lj_$TAG.csv \
    python profiles/prof_driver.py lj > /dev/null 2>&1
ls -la $OUT
