set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_parallel.py -x -q 2>&1 | tail -5
python bench.py --gpus 1 --steps 10 --no-cpu > gpurun_out/scale_1.json 2>gpurun_out/scale.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2>>gpurun_out/scale.err
  fi
done
tail -5 gpurun_out/scale.err
for f in gpurun_out/scale_*.json; do python -c "
import json,sys
for l in open('$f'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print('$f', j['n_gpus'], round(j['value']), j['ms_per_step'], j['spmm']['ms'], j['scoring']['ms'], j['e2e']['value'])
"; done
