export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
cap() { # name regex skip count
  ncu --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c $4 -o gpurun_out/r3n_$1 -f python scripts/quick_time.py 25 100000 > gpurun_out/r3n_$1.log 2>&1
}
cap fin 'k_finish' 3 2
cap pf 'k_pf' 3 1
cap diag 'k_diag_prod' 3 1
ls -la gpurun_out | grep r3n
