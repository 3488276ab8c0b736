# targeted full captures (indices from the launch list of the same build, profiles/NOTES.md)
export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
cap() { # name regex skip count
  ncu --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c $4 -o gpurun_out/p14_$1 -f python scripts/quick_time.py 25 100000 > gpurun_out/p14_$1.log 2>&1
}
cap pf 'k_pf|k_diag_prod' 6 3
