# round-2 end-of-round measurement batch on ONE GPU: parity tests, bench (both arms, n = 25 and n = 20), LUAD (C1), the
# C3 fit, and the ncu captures of the final build.   usage: bash scripts/final_measure_r2.sh [quick]   (quick = no ncu)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.txt 2>&1; tail -3 gpurun_out/r2z_tests.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 400 gpurun_out/r2z_bench_n1.json
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; tail -c 300 gpurun_out/r2z_bench_ref.json
python bench.py --gpus 1 --steps 10 --warmup 3 --n 20 --patients 10000 > gpurun_out/r2z_bench_n20.json 2> gpurun_out/r2z_bench_n20.err; tail -c 300 gpurun_out/r2z_bench_n20.json
python scripts/luad_fit.py > gpurun_out/r2z_luad.json 2> gpurun_out/r2z_luad.err; tail -c 600 gpurun_out/r2z_luad.json
python scripts/c3_fit.py > gpurun_out/r2z_c3.json 2> gpurun_out/r2z_c3.err; tail -c 600 gpurun_out/r2z_c3.json
[ "$1" = quick ] && exit 0
export MMH_GRAPH=0 MMH_STREAMS=1 EVALS=1 VALUE_ONLY=0
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 20000 --csv --log-file gpurun_out/r2z_launches_n25.csv python scripts/quick_time.py 25 100000 > gpurun_out/r2z_ncu_a.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_solve_tile -s 1500 -c 24 -o gpurun_out/r2z_solve_tile -f python scripts/quick_time.py 25 100000 > gpurun_out/r2z_ncu_b.log 2>&1
ls -la gpurun_out | grep r2z
