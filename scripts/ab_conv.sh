set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sg3_ops_gpu.py tests/test_rrdb_gpu.py -x -q -k "modulated_conv2d or rrdb" 2>&1 | tail -3 > gpurun_out/ab_tests.log
python scripts/layer_times.py 16 T > gpurun_out/ab_T_default.txt 2>&1
timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/ab_c5.json 2> gpurun_out/ab_c5.err
cat gpurun_out/ab_tests.log; grep -E "L1[0-3]|total|conv'" gpurun_out/ab_T_default.txt; cut -c1-200 gpurun_out/ab_c5.json; grep -o '"kernel_ms_per_step[^}]*}' gpurun_out/ab_c5.json
