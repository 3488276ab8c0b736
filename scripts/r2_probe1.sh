# round-2 probe of the first blocked-solve build: launch list, one full capture, chunk sweep
export EVALS=1 VALUE_ONLY=0
MMH_GRAPH=0 MMH_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python scripts/quick_time.py 25 100000 > gpurun_out/r2a_ncu_a.log 2>&1
MMH_GRAPH=0 MMH_STREAMS=1 ncu --set full --import-source on --clock-control none -k regex:k_blk -s 60 -c 8 -o gpurun_out/r2a_blk -f python scripts/quick_time.py 25 100000 > gpurun_out/r2a_ncu_b.log 2>&1
python scripts/sweep_chunks.py 25 100000 1073741824,4294967296,8589934592 6 > gpurun_out/r2a_sweep.txt 2>&1
cat gpurun_out/r2a_sweep.txt
ls -la gpurun_out | tail -5
