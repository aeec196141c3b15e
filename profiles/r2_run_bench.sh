#!/bin/bash
# GPU-box driver: bench (no profiler), then the ncu launch list with DRAM traffic, one --set full capture of the dominant kernel, other configs
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_p_bench.json 2> gpurun_out/r2_p_bench.err; echo "bench rc $?"; cut -c1-400 gpurun_out/r2_p_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc $?"
python profiles/ncu_traffic.py gpurun_out/r2_bench_launches_ncu.csv 364 gpurun_out/r2_tc2_traffic.json | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_i16_tc2 -s 1 -c 1 -o gpurun_out/r2_tc2_persist -f python profiles/run_tc2_layer.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
Y2_SIZE=608 timeout 600 python profiles/layer_table.py 128 > gpurun_out/r2_p_lt_608.json 2> gpurun_out/r2_p_lt_608.err; echo "608 rc $?"
Y2_CLASSES=20 timeout 600 python profiles/layer_table.py 64 > gpurun_out/r2_p_lt_voc.json 2> gpurun_out/r2_p_lt_voc.err; echo "voc rc $?"
python profiles/lt_print.py gpurun_out/r2_p_lt_608.json; python profiles/lt_print.py gpurun_out/r2_p_lt_voc.json | head -2
