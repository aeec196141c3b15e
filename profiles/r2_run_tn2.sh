#!/bin/bash
# GPU-box driver: rounding-group variants after a change of csrc/conv_i16_tc32.cu: parity tests, layer tables Tn = 32 / 16 / 8, timeline
cd /root/repo; V=$PWD/yolo-fpga-accelerator_b200/lib/variants; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_width.py -x -q -m gpu -k "tn8_tn16 or rounding_group or tn32 or tile_param or tn_variant" > gpurun_out/tn_tests.log 2>&1; echo "tn tests rc $?"; tail -3 gpurun_out/tn_tests.log
for tn in 32 16 8; do
  b=32; [ $tn = 32 ] && b=64
  Y2_TN=$tn timeout 200 python profiles/layer_table.py $b > gpurun_out/lt2_tn$tn.json 2> gpurun_out/lt2_tn$tn.err; echo "tn$tn rc $?"
  python profiles/lt_print.py gpurun_out/lt2_tn$tn.json | grep -E "fps|^(4|8|12|13|19|23|29) "
done
for tn in 32 16; do YOLO2CUDA_LIB=$V/libyolo2cuda_prof32.so Y2_TN=$tn Y2_REPS=2 timeout 120 python profiles/tc32_timeline.py > gpurun_out/tc32_tl2_tn$tn.txt 2>&1; echo "timeline rc $?"; done
