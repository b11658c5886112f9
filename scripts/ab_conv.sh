set -x
mkdir -p gpurun_out
for v in "" nohint hint1us; do
  if [ -n "$v" ]; then export MAUA_B200_LIB=$PWD/maua_b200/lib/libmaua_$v.so; else unset MAUA_B200_LIB; fi
  python scripts/layer_times.py 16 T > gpurun_out/ab_T_pms_$v.txt 2>&1
  MBOPT_CONV_CM_STACK=1 timeout 300 python scripts/layer_times.py 16 T > gpurun_out/ab_T_cms_$v.txt 2>&1
done
grep -E "L1[123]|total|conv'" gpurun_out/ab_T_pms_*.txt gpurun_out/ab_T_cms_*.txt
