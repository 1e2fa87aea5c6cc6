set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q 2>&1 | tail -5
python bench.py --no-cpu > gpurun_out/bench12_v2.json 2>gpurun_out/bench12.err; python -c "
import json; j=json.load(open('gpurun_out/bench12_v2.json')); print('v2', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
LGX_SCORE_EPILOGUE=1 python bench.py --no-cpu > gpurun_out/bench12_v1.json 2>>gpurun_out/bench12.err; python -c "
import json; j=json.load(open('gpurun_out/bench12_v1.json')); print('v1', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
python bench.py --no-cpu --mode bf16x3 > gpurun_out/bench12_v2x3.json 2>>gpurun_out/bench12.err; python -c "
import json; j=json.load(open('gpurun_out/bench12_v2x3.json')); print('v2x3', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
tail -3 gpurun_out/bench12.err
