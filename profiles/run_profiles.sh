#!/bin/bash
# Run on the GPU box (gpurun -- bash profiles/run_profiles.sh <tag>): launch lists + one full-set capture
# per dominant kernel.  Outputs land in gpurun_out/; `python profiles/summarize.py <tag>` then writes the tracked
# summaries under profiles/.
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
# (1) launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/bench_under_ncu_$TAG.log 2>&1
# (2) full-set captures: gravity (Newton's-third-law all-pairs), LJ over Verlet lists, the fused LJ step (option),
#     LJ cell scan (the slab path), the list build
ncu --set full --clock-control none --import-source on -k regex:sym_kernel -s 1 -c 1 -f -o $OUT/prof_gravity_$TAG \
    python profiles/prof_driver.py gravity > $OUT/prof_gravity_$TAG.log 2>&1
PROF_STEPS=4 ncu --set full --clock-control none --import-source on -k regex:verlet_force -s 2 -c 1 -f -o $OUT/prof_ljverlet_$TAG \
    python profiles/prof_driver.py lj > $OUT/prof_ljverlet_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fz_step -s 6 -c 1 -f -o $OUT/prof_ljfused_$TAG \
    python profiles/prof_fused.py 1 > $OUT/prof_ljfused_$TAG.log 2>&1
PROF_STEPS=4 PROF_VERLET=0 ncu --set full --clock-control none --import-source on -k regex:cell_pairs2 -s 2 -c 1 -f -o $OUT/prof_ljscan_$TAG \
    python profiles/prof_driver.py lj > $OUT/prof_ljscan_$TAG.log 2>&1
# (3) launch lists of the LJ and the water step (eager launches of nbx_step_vv)
PROF_STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_lj_$TAG.csv \
    python profiles/prof_driver.py lj > /dev/null 2>&1
PROF_STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_water_$TAG.csv \
    python profiles/prof_driver.py water > /dev/null 2>&1
ls -la $OUT | tail -20
