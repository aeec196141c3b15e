#!/bin/bash
# usage: r2_run_tc32_variants.sh name... : Tn = 32 (64 frames) and Tn = 16 (32 frames) layer tables of the default library and of
# each lib/variants/libyolo2cuda_<name>.so (profiles/build_variant_tc32.sh)
cd /root/repo; V=$PWD/yolo-fpga-accelerator_b200/lib/variants; mkdir -p gpurun_out
for v in default "$@"; do
  if [ $v = default ]; then unset YOLO2CUDA_LIB; else export YOLO2CUDA_LIB=$V/libyolo2cuda_$v.so; fi
  for tn in ${Y2_TNS:-32 16}; do
    b=32; [ $tn = 32 ] && b=64
    Y2_TN=$tn timeout 200 python profiles/layer_table.py $b > gpurun_out/t32_${v}_tn$tn.json 2> gpurun_out/t32_${v}_tn$tn.err; echo "$v tn$tn rc $?"
    python profiles/lt_print.py gpurun_out/t32_${v}_tn$tn.json | grep -E "fps|^(4|8|12|19|23|29) "
  done
done
