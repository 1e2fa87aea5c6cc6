set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench8_switch.json 2>gpurun_out/bench8.err; python -c "
import json; j=json.load(open('gpurun_out/bench8_switch.json')); print('switch', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
LGX_LIB_PATH=$PWD/factors_of_serendipity_recommendation_b200/liblgx_alt.so python bench.py --no-cpu > gpurun_out/bench8_sweep.json 2>>gpurun_out/bench8.err; python -c "
import json; j=json.load(open('gpurun_out/bench8_sweep.json')); print('sweep', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
python bench.py --no-cpu --mode bf16x3 > gpurun_out/bench8_x3.json 2>>gpurun_out/bench8.err; python -c "
import json; j=json.load(open('gpurun_out/bench8_x3.json')); print('x3', j['value'], j['spmm']['ms'], j['scoring'], j['e2e']['value'])"
