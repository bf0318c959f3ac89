set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --windows 8 --steps 1 --warmup 1 --cpu-windows 1 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --windows 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_r1.csv
