#!/bin/bash
# GPU-box driver: Tn = 8 / 16 tensor-core kernel (csrc/conv_i16_tc32.cu, TNW template): parity tests, then per-layer tables with the
# tensor-core kernel forced on every eligible layer (YOLO2CUDA_TC=2), off (=0) and the default policy
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_width.py -x -q -m gpu -k "tn8_tn16 or rounding_group or tn32 or tile_param or tn_variant" > gpurun_out/tn_tests.log 2>&1; echo "tn tests rc $?"; tail -5 gpurun_out/tn_tests.log
for tn in 16 8; do
  for tc in 2 0 auto; do
    if [ $tc = auto ]; then unset YOLO2CUDA_TC; else export YOLO2CUDA_TC=$tc; fi
    Y2_TN=$tn timeout 300 python profiles/layer_table.py 32 > gpurun_out/lt_tn${tn}_tc${tc}.json 2> gpurun_out/lt_tn${tn}_tc${tc}.err; echo "lt tn$tn tc$tc rc $?"
  done
done
unset YOLO2CUDA_TC
Y2_TN=32 timeout 300 python profiles/layer_table.py 64 > gpurun_out/lt_tn32.json 2> gpurun_out/lt_tn32.err; echo "lt tn32 rc $?"
