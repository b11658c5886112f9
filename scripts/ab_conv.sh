set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sg3_ops_gpu.py -x -q -k modulated_conv2d 2>&1 | tail -5 > gpurun_out/ab_tests.log
python scripts/layer_times.py 16 T > gpurun_out/ab_T_default.txt 2>&1
python scripts/layer_times.py 16 R > gpurun_out/ab_R_default.txt 2>&1
cat gpurun_out/ab_tests.log; grep -E "L|total|conv'" gpurun_out/ab_T_default.txt; grep -E "L1[0-3]|total|conv'" gpurun_out/ab_R_default.txt
