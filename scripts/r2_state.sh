# Round-2 state check on one B200: GPU tests, smoke, per-layer table
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
python scripts/layer_times.py 16 T > gpurun_out/r2s_layers_T16.txt 2>&1
tail -3 gpurun_out/r2s_tests.log; cat gpurun_out/r2s_smoke.log; grep -E "L1[123]|total|conv'" gpurun_out/r2s_layers_T16.txt
