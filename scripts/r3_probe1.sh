export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
ncu --set full --import-source on --clock-control none -k regex:k_solve_rb -s 5 -c 3 -o gpurun_out/r3b_rb_fwd -f python scripts/quick_time.py 25 100000 > gpurun_out/r3b_ncu_b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_solve_rb -s 26 -c 3 -o gpurun_out/r3b_rb_adj -f python scripts/quick_time.py 25 100000 > gpurun_out/r3b_ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_solve_rb -s 15 -c 3 -o gpurun_out/r3b_rb_thin -f python scripts/quick_time.py 25 100000 > gpurun_out/r3b_ncu_d.log 2>&1
ls -la gpurun_out | tail -4
