# round-end measurement batch (one GPU): parity tests, bench (both arms), LUAD, ncu captures of the final build
# usage: bash scripts/final_measure.sh [quick]     (quick = no ncu passes)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
python scripts/luad_fit.py > gpurun_out/luad.json 2>&1; cat gpurun_out/luad.json
python scripts/prof_classes.py 20 10000 > gpurun_out/n20.txt 2>&1; cat gpurun_out/n20.txt
[ "$1" = quick ] && exit 0
export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
ncu --set full --import-source on --clock-control none -k regex:k_solve_tile -s 1500 -c 24 -o gpurun_out/r1_final_solve_tile -f python scripts/quick_time.py 25 100000 > gpurun_out/ncu_final_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python scripts/quick_time.py 25 100000 > gpurun_out/ncu_final_b.log 2>&1
ls -la gpurun_out | tail -12
