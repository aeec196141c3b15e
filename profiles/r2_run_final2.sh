#!/bin/bash
# GPU-box driver, end of round 2 (tight timeouts: 4.5 GPU-minutes were left): rounding-group variant tests first (the six-builder
# conv_i16_tc32 kernel), then the rest of the GPU suite, smoke, bench, layer tables of the Tn = 32 / 16 / 8 variants
cd /root/repo; mkdir -p gpurun_out
K="tn8_tn16 or rounding_group or tn32 or tile_param or tn_variant"
timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_width.py -x -q -m gpu -k "$K" > gpurun_out/f2_tn_tests.log 2>&1; rc=$?; echo "tn tests rc $rc"; tail -2 gpurun_out/f2_tn_tests.log
if [ $rc != 0 ]; then exit 1; fi
timeout 170 python -m pytest tests -x -q -m gpu -k "not ($K)" > gpurun_out/f2_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/f2_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/f2_smoke.log
timeout 100 python bench.py --steps 5 --warmup 3 > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc $?"; cut -c1-160 gpurun_out/f2_bench.json
for tn in 32 16 8; do
  b=32; [ $tn = 32 ] && b=64
  Y2_TN=$tn timeout 40 python profiles/layer_table.py $b > gpurun_out/f2_lt_tn$tn.json 2> gpurun_out/f2_lt_tn$tn.err; echo "tn$tn rc $?"
  python profiles/lt_print.py gpurun_out/f2_lt_tn$tn.json 2>/dev/null | grep -E "fps|^(8|19|23) "
done
