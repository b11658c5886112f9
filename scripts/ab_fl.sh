mkdir -p gpurun_out
for v in "" ring6 ring4 "" ring6 ring4; do
  if [ -n "$v" ]; then export MAUA_B200_LIB=$PWD/maua_b200/lib/libmaua_$v.so; else unset MAUA_B200_LIB; fi
  echo "variant: ${v:-default}"; python scripts/layer_times.py 16 T 2>&1 | grep -E "L10|L11|styles"
done
