set -x
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'flrelu|conv_' --launch-skip 28 --launch-count 28 --csv --log-file gpurun_out/traffic_T_b16.csv python scripts/one_forward.py 16 T > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'flrelu|conv_' --launch-skip 28 --launch-count 28 --csv --log-file gpurun_out/traffic_R_b16.csv python scripts/one_forward.py 16 R > /dev/null 2>&1
python -m pytest tests/test_video_writer_gpu.py tests/test_render_entry_gpu.py tests/test_sample_generate_gpu.py tests/test_maua_alias_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2t_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
cat gpurun_out/r2t_tests.log gpurun_out/r2t_bench.json
