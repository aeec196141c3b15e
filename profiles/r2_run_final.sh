#!/bin/bash
# GPU-box driver: deadlock-detecting build, parity with 3 persistent CTAs, full GPU suite, bench + ncu launch list with DRAM traffic
cd /root/repo; V=$PWD/yolo-fpga-accelerator_b200/lib/variants; mkdir -p gpurun_out
YOLO2CUDA_LIB=$V/libyolo2cuda_prof5.so YOLO2CUDA_TC=2 Y2_REPS=1 Y2_SHAPES="64,128,3,104,104;512,256,3,13,13;1024,512,1,13,13;24,130,3,13,13" timeout 300 python profiles/tc2_role_profile.py > gpurun_out/f_prof5.txt 2>&1
echo "prof5 rc $? deadlocks $(grep -c DEADLOCK gpurun_out/f_prof5.txt)"
if grep -q DEADLOCK gpurun_out/f_prof5.txt; then exit 1; fi
YOLO2CUDA_LIB=$V/libyolo2cuda_grid3.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core" > gpurun_out/f_grid3_tests.log 2>&1; echo "grid3 rc $?"; tail -2 gpurun_out/f_grid3_tests.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/f_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"; cut -c1-300 gpurun_out/f_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc $?"
python profiles/ncu_traffic.py gpurun_out/r2_bench_launches_ncu.csv 735 gpurun_out/r2_tc2_traffic.json
