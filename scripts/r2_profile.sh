#!/bin/bash
# Round-2 ncu evidence (one GPU).  Launch list of the bench command + full captures of the dominant kernels.
set -x
O=gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline --no-k26 > $O/r2_launches_bench.json 2> $O/r2_launches.err
# k_walk3 on a converged iteration (predicted gathers) and on iteration 2 (wandering colony: neighbour prefetch, four-entry probes)
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk3 -s 60 -c 1 -f -o $O/r2_walk3_converged python scripts/profile_run.py 70 4 > $O/r2_walk3_converged.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk3 -s 1 -c 1 -f -o $O/r2_walk3_wandering python scripts/profile_run.py 3 4 > $O/r2_walk3_wandering.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk_batch3 -s 4 -c 1 -f -o $O/r2_walk_batch3 python scripts/profile_batch.py 256 512 6 > $O/r2_walk_batch3.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk26p -s 4 -c 1 -f -o $O/r2_walk26p python scripts/profile_run.py 6 4 26 > $O/r2_walk26p.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none -k regex:k_evaporate_tiles -s 60 -c 1 -f -o $O/r2_evaporate python scripts/profile_run.py 70 4 > $O/r2_evaporate.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_batch.csv python scripts/profile_batch.py 256 512 6 > $O/r2_launches_batch.log 2>&1
ls -la $O/*.ncu-rep
