cd /root/repo
bash profiles/r2_run_variants.sh r104 r104ld2 2>&1 | grep -E "rc|fps|^(4|8|12|19|23|29) "
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pass_schedule or tensor_core_conv" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('gpurun_out/g_bench.json')); print(d['value'], d['e2e']['value'], d['config']['workload'], d['parity']['mismatches'], d['roofline']['traffic'])"
