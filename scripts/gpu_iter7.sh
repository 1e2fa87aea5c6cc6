set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q 2>&1 | tail -3
for m in bf16 bf16x3; do python bench.py --no-cpu --mode $m > gpurun_out/bench7_$m.json 2>gpurun_out/bench7.err; python -c "
import json; j=json.load(open('gpurun_out/bench7_$m.json')); print(j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"; done
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_score_topk_tc' -s 3 -c 1 -o gpurun_out/prof_tc_r1e python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log | cut -c1-200
