# serialised launch list of one evaluation: usage ncu_launchlist.sh n patients out.csv
export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 20000 --csv --log-file $3 python scripts/quick_time.py $1 $2 > $3.log 2>&1
