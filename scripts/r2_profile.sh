#!/bin/bash
# Round-2 ncu evidence (one GPU).  Launch list of the bench command + full captures of the dominant kernels.
set -x
O=gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline --no-k26 > $O/r2_launches_bench.json 2> $O/r2_launches.err
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk2 -s 120 -c 1 -f -o $O/r2_walk2 python scripts/profile_run.py 70 4 > $O/r2_walk2.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_walk_batch -s 4 -c 1 -f -o $O/r2_walk_batch python scripts/profile_batch.py 256 512 6 > $O/r2_walk_batch.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_batch_deposit -s 3 -c 1 -f -o $O/r2_batch_deposit python scripts/profile_batch.py 256 512 6 > $O/r2_batch_deposit.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_gtsp_iterate -s 1 -c 1 -f -o $O/r2_gtsp python scripts/gtsp_bench.py 256 128 2 > $O/r2_gtsp.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_batch.csv python scripts/profile_batch.py 256 512 6 > $O/r2_launches_batch.log 2>&1
ls -la $O/*.ncu-rep
