#!/bin/bash
# usage: r2_run_variants.sh name... : 126-frame layer table of the default library and of each lib/variants/libyolo2cuda_<name>.so
cd /root/repo; V=$PWD/yolo-fpga-accelerator_b200/lib/variants; mkdir -p gpurun_out
N=${Y2_FRAMES:-126}
timeout 300 python profiles/layer_table.py $N > gpurun_out/v_default.json 2> gpurun_out/v_default.err; echo "default rc $?"; python profiles/lt_print.py gpurun_out/v_default.json | grep -E "fps|^(2|4|5|8|12|19|23|29) "
for v in "$@"; do
  YOLO2CUDA_LIB=$V/libyolo2cuda_$v.so timeout 300 python profiles/layer_table.py $N > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err; echo "$v rc $?"; python profiles/lt_print.py gpurun_out/v_$v.json | grep -E "fps|^(2|4|5|8|12|19|23|29) "
done
