#!/bin/bash
cd /root/repo; V=$PWD/yolo-fpga-accelerator_b200/lib/variants; mkdir -p gpurun_out
YOLO2CUDA_LIB=$V/libyolo2cuda_prof5.so YOLO2CUDA_TC=2 Y2_REPS=1 Y2_SHAPES="32,64,3,52,52;128,64,1,26,26;64,128,3,104,104;96,40,1,19,19" timeout 300 python profiles/tc2_role_profile.py > gpurun_out/h_prof5.txt 2>&1
echo "prof5 rc $? deadlocks $(grep -c DEADLOCK gpurun_out/h_prof5.txt)"; grep -E "^[0-9]+ [0-9]+ [0-9]" gpurun_out/h_prof5.txt; grep -i "error" gpurun_out/h_prof5.txt | head -3
if grep -q DEADLOCK gpurun_out/h_prof5.txt; then grep -A26 DEADLOCK gpurun_out/h_prof5.txt | head -40; exit 1; fi
timeout 120 python profiles/tc2_one_case.py 32,64,3,52,52 128,64,1,26,26 36,64,3,13,13 96,40,1,19,19 17,33,3,20,11 64,128,3,26,26 2>&1 | tail -7
YOLO2CUDA_LIB=$V/libyolo2cuda_grid3.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core" > gpurun_out/h_grid3_tests.log 2>&1; echo "grid3 rc $?"; tail -3 gpurun_out/h_grid3_tests.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/h_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/h_tests.log
timeout 300 python profiles/layer_table.py 126 > gpurun_out/h_lt.json 2> gpurun_out/h_lt.err; echo "lt rc $?"; python profiles/lt_print.py gpurun_out/h_lt.json
