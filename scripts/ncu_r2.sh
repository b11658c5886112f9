set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'conv_pms' --launch-skip 3 --launch-count 3 -o gpurun_out/r2_pms -f python scripts/one_forward.py 16 T > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
ls -la gpurun_out/*.ncu-rep
