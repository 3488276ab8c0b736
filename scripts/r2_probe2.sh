export EVALS=1 VALUE_ONLY=0
MMH_GRAPH=0 MMH_STREAMS=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_blk --csv --log-file gpurun_out/r2b_launches.csv python scripts/quick_time.py 25 100000 > gpurun_out/r2b_ncu_a.log 2>&1
MMH_GRAPH=0 MMH_STREAMS=1 ncu --set full --import-source on --clock-control none -k regex:k_blk -s 60 -c 6 -o gpurun_out/r2b_blk -f python scripts/quick_time.py 25 100000 > gpurun_out/r2b_ncu_b.log 2>&1
ls -la gpurun_out | tail -4
