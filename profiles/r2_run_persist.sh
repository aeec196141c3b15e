#!/bin/bash
# GPU-box driver for the persistent + TMA tcgen05 conv: deadlock-detecting build first, then parity, then timing
cd /root/repo
V=yolo-fpga-accelerator_b200/lib/variants
mkdir -p gpurun_out
echo "== prof5 (deadlock-detecting, 5 persistent CTAs)"
YOLO2CUDA_LIB=$PWD/$V/libyolo2cuda_prof5.so YOLO2CUDA_TC=2 Y2_REPS=1 Y2_SHAPES="64,128,3,104,104;512,256,3,13,13;1024,512,1,13,13;24,130,3,13,13" timeout 300 python profiles/tc2_role_profile.py > gpurun_out/p1_prof5.txt 2>&1
echo "rc $?"; grep -c DEADLOCK gpurun_out/p1_prof5.txt; grep -E "^[0-9]+ [0-9]+ [0-9]" gpurun_out/p1_prof5.txt
if grep -q DEADLOCK gpurun_out/p1_prof5.txt; then grep -A26 DEADLOCK gpurun_out/p1_prof5.txt | head -60; exit 1; fi
echo "== grid3 parity (tensor-core tests, 3 persistent CTAs)"
YOLO2CUDA_LIB=$PWD/$V/libyolo2cuda_grid3.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core" > gpurun_out/p1_grid3_tests.log 2>&1
echo "rc $?"; tail -5 gpurun_out/p1_grid3_tests.log
echo "== default lib, all gpu tests"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p1_tests.log 2>&1
echo "rc $?"; tail -5 gpurun_out/p1_tests.log
echo "== layer tables"
timeout 300 python profiles/layer_table.py 128 > gpurun_out/p1_lt_default.json 2> gpurun_out/p1_lt_default.err; echo "rc $?"
YOLO2CUDA_TC=2 timeout 300 python profiles/layer_table.py 128 > gpurun_out/p1_lt_tc2.json 2> gpurun_out/p1_lt_tc2.err; echo "rc $?"
python profiles/lt_print.py gpurun_out/p1_lt_default.json
python profiles/lt_print.py gpurun_out/p1_lt_tc2.json
