set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py tests/test_gpu_propagate.py -x -q 2>&1 | tail -5
for v in 0 1 2 3 4; do LGX_SPMM_VARIANT=$v python scripts/spmm_sweep.py amazon-book; done 2>&1 | grep variant | tee gpurun_out/spmm_sweep.jsonl
for v in 5 6 7; do for h in 64 256 1024; do LGX_SPMM_VARIANT=$v LGX_SPMM_HOT_DEGREE=$h python scripts/spmm_sweep.py amazon-book; done; done 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep.jsonl
for c in 64 128 512 1024; do LGX_CHUNK=$c python scripts/spmm_sweep.py amazon-book; done 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep.jsonl
python bench.py --no-cpu --steps 10 > gpurun_out/bench2_bf16.json 2>gpurun_out/bench2.err; python -c "
import json; j=json.load(open('gpurun_out/bench2_bf16.json')); print(j['value'], j['spmm'], j['scoring'], j['e2e'])"
python bench.py --no-cpu --steps 10 --mode bf16x3 > gpurun_out/bench2_bf16x3.json 2>>gpurun_out/bench2.err; python -c "
import json; j=json.load(open('gpurun_out/bench2_bf16x3.json')); print(j['value'], j['spmm'], j['scoring'], j['e2e'])"
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_score_topk_tc' -s 3 -c 1 -o gpurun_out/prof_tc_r1 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
