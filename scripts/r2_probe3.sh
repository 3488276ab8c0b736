export EVALS=1 VALUE_ONLY=0
MMH_GRAPH=0 MMH_STREAMS=1 ncu --set full --import-source on --clock-control none -k regex:k_blk -s 62 -c 2 -o gpurun_out/r2c_blk -f python scripts/quick_time.py 25 100000 > gpurun_out/r2c_ncu_b.log 2>&1
ls -la gpurun_out | tail -3
