set -x
mkdir -p gpurun_out
MBOPT_CONV_CM_STACK=1 ncu --set full --clock-control none --import-source on -k regex:'conv_cms' --launch-skip 2 --launch-count 2 -o gpurun_out/r2_cms -f python scripts/one_forward.py 16 T > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
ls -la gpurun_out/r2_cms.ncu-rep
