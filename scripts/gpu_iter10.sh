set -x
mkdir -p gpurun_out
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; sleep 2; done ) > gpurun_out/mem_1b.log 2>&1 &
MON=$!
timeout 900 python bench.py --no-cpu --workload synth-1b --steps 3 --warmup 3 > gpurun_out/bench10_1b_n1.json 2>gpurun_out/bench10.err
echo rc=$?
kill $MON
sort -n gpurun_out/mem_1b.log | tail -1
python -c "
import json; j=json.load(open('gpurun_out/bench10_1b_n1.json')); print('1b', j['value'], j['ms_per_step'], j['spmm'], j['scoring'], j['roofline_spmm'])"
tail -5 gpurun_out/bench10.err
