set -x
mkdir -p gpurun_out
python bench.py --no-cpu > gpurun_out/bench9.json 2>gpurun_out/bench9.err; python -c "
import json; j=json.load(open('gpurun_out/bench9.json')); print('amazon', j['value'], j['spmm'], j['scoring'], j['e2e']['value'])"
python bench.py --no-cpu --workload synth-10m --steps 10 > gpurun_out/bench9_10m.json 2>>gpurun_out/bench9.err; python -c "
import json; j=json.load(open('gpurun_out/bench9_10m.json')); print('10m', j['value'], j['spmm'], j['scoring'], j['e2e']['value'])"
python bench.py --no-cpu --workload synth-100m --steps 5 --warmup 3 > gpurun_out/bench9_100m.json 2>>gpurun_out/bench9.err; python -c "
import json; j=json.load(open('gpurun_out/bench9_100m.json')); print('100m', j['value'], j['spmm'], j['scoring'], j['e2e']['value'], j['roofline_spmm'])"
nvidia-smi --query-gpu=memory.used --format=csv
tail -5 gpurun_out/bench9.err
