set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sg3_ops_gpu.py -x -q -k "modulated_conv2d and 12" 2>&1 | tail -5 > gpurun_out/ab_tests.log
MB_RRDB_CMS=1 timeout 300 python -m pytest tests/test_rrdb_gpu.py -x -q -s 2>&1 | tail -8 >> gpurun_out/ab_tests.log
python scripts/layer_times.py 16 T > gpurun_out/ab_T_default.txt 2>&1
MBOPT_CONV_CM_STACK=1 timeout 300 python scripts/layer_times.py 16 T > gpurun_out/ab_T_cms.txt 2>&1
timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/ab_c5_pms.json 2> gpurun_out/ab_c5_pms.err
MB_RRDB_CMS=1 timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/ab_c5_cms.json 2> gpurun_out/ab_c5_cms.err
cat gpurun_out/ab_tests.log; grep -E "L1[0-3]|total|conv'" gpurun_out/ab_T_default.txt gpurun_out/ab_T_cms.txt; for f in pms cms; do cut -c1-160 gpurun_out/ab_c5_$f.json; grep -o '"kernel_ms_per_step[^}]*}' gpurun_out/ab_c5_$f.json; done; tail -2 gpurun_out/ab_c5_cms.err
