#!/bin/bash
# GPU-box driver: records of the final build - ncu --set full of the tcgen05 kernel (128-channel and HALF instantiations), the other BASELINE configs, fp32 line
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_i16_tc2 -s 1 -c 1 -o gpurun_out/r2_tc2_final -f python profiles/run_tc2_layer.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_i16_tc2 -s 1 -c 1 -o gpurun_out/r2_tc2_final_half -f python profiles/run_tc2_layer.py 32 64 3 208 208 > gpurun_out/ncu_full_half.log 2>&1; echo "ncu half rc $?"
Y2_SIZE=608 timeout 600 python profiles/layer_table.py 128 > gpurun_out/r2_f_lt_608.json 2> gpurun_out/r2_f_lt_608.err; echo "608 rc $?"
Y2_CLASSES=20 timeout 600 python profiles/layer_table.py 64 > gpurun_out/r2_f_lt_voc.json 2> gpurun_out/r2_f_lt_voc.err; echo "voc rc $?"
timeout 600 python profiles/layer_table.py 126 > gpurun_out/r2_f_lt_416.json 2> gpurun_out/r2_f_lt_416.err; echo "416 rc $?"
python profiles/lt_print.py gpurun_out/r2_f_lt_608.json | head -1; python profiles/lt_print.py gpurun_out/r2_f_lt_voc.json | head -1; python profiles/lt_print.py gpurun_out/r2_f_lt_416.json
timeout 900 python bench.py --precision fp32 --frames-per-gpu 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_f_bench_fp32.json 2> gpurun_out/r2_f_bench_fp32.err; echo "fp32 rc $?"; cut -c1-200 gpurun_out/r2_f_bench_fp32.json
