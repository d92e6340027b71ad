ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 1 --warmup 3 --resident-only --hot-only --no-cpu-baseline --no-profile --no-graph > gpurun_out/ncu_final.log 2>&1
wc -l gpurun_out/r02_ncu_launches_final.csv
