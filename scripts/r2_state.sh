# Round-2 state check on one B200: GPU tests, smoke, bench lines (c2 headline, c3, sg2), per-layer tables
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2s_bench_c3.json 2> gpurun_out/r2s_bench_c3.err
python bench.py --config sg2 --no-cpu-baseline > gpurun_out/r2s_bench_sg2.json 2> gpurun_out/r2s_bench_sg2.err
tail -3 gpurun_out/r2s_tests.log; cat gpurun_out/r2s_smoke.log gpurun_out/r2s_bench.json gpurun_out/r2s_bench_c3.json
