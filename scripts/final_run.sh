set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s5_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s5_bench_ref.json 2> gpurun_out/s5_bench_ref.err
python bench.py > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s5_ncu_bench.log 2>&1
cat gpurun_out/s5_tests.log gpurun_out/s5_smoke.log gpurun_out/s5_bench.json
