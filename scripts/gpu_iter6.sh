set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_train.py --epochs 2 > gpurun_out/train_yelp_fused.json 2>gpurun_out/train.err; tail -c 1500 gpurun_out/train_yelp_fused.json
python scripts/bench_train.py --epochs 2 --fused-adam 0 --cpu-steps 0 > gpurun_out/train_yelp_torchadam.json 2>>gpurun_out/train.err; tail -c 600 gpurun_out/train_yelp_torchadam.json
python bench.py --no-cpu > gpurun_out/bench6.json 2>gpurun_out/bench6.err; python -c "
import json; j=json.load(open('gpurun_out/bench6.json')); print(j['value'], j['ms_per_step'], j['spmm'], j['scoring'], j['e2e'], j['clocks'])"
