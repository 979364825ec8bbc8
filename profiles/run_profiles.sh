#!/bin/bash
# Run on the GPU box (gpurun -- bash profiles/run_profiles.sh <tag>): launch lists + one full-set capture
# per dominant kernel.  Outputs land in gpurun_out/; `python profiles/summarize.py <tag>` then writes the tracked
# summaries under profiles/.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
# (1) launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > $OUT/bench_under_ncu_$TAG.log 2>&1
# (2) full-set captures: gravity (Newton's-third-law all-pairs), LJ over Verlet lists (1,048,576 atoms and one rank's share
#     of the 8-GPU run), the periodic Coulomb cutoff kernel of water (cutoff 0.49 L), the dipole kernel
ncu --set full --clock-control none --import-source on -k regex:sym_kernel -s 1 -c 1 -f -o $OUT/prof_gravity_$TAG \
    python profiles/prof_driver.py gravity > $OUT/prof_gravity_$TAG.log 2>&1
PROF_STEPS=4 ncu --set full --clock-control none --import-source on -k regex:verlet_force -s 2 -c 1 -f -o $OUT/prof_ljverlet_$TAG \
    python profiles/prof_driver.py lj > $OUT/prof_ljverlet_$TAG.log 2>&1
PROF_STEPS=4 ncu --set full --clock-control none --import-source on -k regex:verlet_force -s 2 -c 1 -f -o $OUT/prof_ljverlet131k_$TAG \
    python profiles/prof_driver.py lj 32 > $OUT/prof_ljverlet131k_$TAG.log 2>&1
PROF_REL=0 PROF_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:sym_kernel -s 0 -c 1 -f -o $OUT/prof_waterpbc_$TAG \
    python profiles/prof_driver.py water > $OUT/prof_waterpbc_$TAG.log 2>&1
PROF_STEPS=1 ncu --set full --clock-control none --import-source on -k regex:allpairs_kernel -s 0 -c 1 -f -o $OUT/prof_dipole_$TAG \
    python profiles/prof_driver.py dipole > $OUT/prof_dipole_$TAG.log 2>&1
# (3) launch lists: the LJ step on one GPU (eager launches of nbx_step_vv), one rank's share of the 8-GPU slab step as a
#     one-slab group (graph replay: the kernel nodes of the regular step), the water step
PROF_STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_lj_$TAG.csv \
    python profiles/prof_driver.py lj > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file $OUT/launches_slab_$TAG.csv \
    python profiles/prof_slab.py 32 40 graph_if_nodes=0 > /dev/null 2>&1
PROF_STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_water_$TAG.csv \
    python profiles/prof_driver.py water > /dev/null 2>&1
ls -la $OUT | tail -20
