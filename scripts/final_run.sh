# Round-end measurement pass on one B200 (run through gpurun): GPU tests, smoke, bench (both arms), launch list.
# DRAM traffic per step (profiles/traffic.json) comes from:
#   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'flrelu|conv_' \
#       --launch-skip 28 --launch-count 28 --csv --log-file gpurun_out/traffic_b16.csv python scripts/one_forward.py 16 T
#   python scripts/make_traffic.py gpurun_out/traffic_b16.csv profiles/traffic.json 16
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s7_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s7_smoke.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s7_bench_ref.json 2> gpurun_out/s7_bench_ref.err
python bench.py > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s7_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s7_ncu_bench.log 2>&1
cat gpurun_out/s7_tests.log gpurun_out/s7_smoke.log gpurun_out/s7_bench.json
