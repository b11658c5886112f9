# Round-2 evidence pass on one B200: per-layer tables, bench lines, ncu launch list, ncu --set full of the top kernels
set -x
mkdir -p gpurun_out

python scripts/e2e_probe.py 720 > gpurun_out/e2e_probe.txt 2>&1
python scripts/layer_times.py 16 T > gpurun_out/r2_layers_T16.txt 2>&1
python scripts/layer_times.py 16 R > gpurun_out/r2_layers_R16.txt 2>&1
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
python bench.py --config sg2 --no-cpu-baseline > gpurun_out/r2_bench_sg2.json 2> gpurun_out/r2_bench_sg2.err
python bench.py --config c5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
# launches 13..27 of the second forward's conv / filter kernels: L6 conv, L6 filter, L7 conv, ..., L13 conv
ncu --set full --clock-control none --import-source on -k regex:'conv_|flrelu_' --launch-skip 41 --launch-count 15 -o gpurun_out/r2_full_b16 -f python scripts/one_forward.py 16 T > gpurun_out/ncu_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/e2e_probe.txt gpurun_out/r2_bench.json
