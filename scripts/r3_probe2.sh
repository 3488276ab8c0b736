export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
ncu --set full --import-source on --clock-control none -k regex:k_solve_rb -s 5 -c 2 -o gpurun_out/r3e_rb_fwd -f python scripts/quick_time.py 25 100000 > gpurun_out/r3e_ncu_b.log 2>&1
ls -la gpurun_out | tail -2
