# usage: gpu_scale.sh N workload [steps]   -- one bench.py run at N GPUs (torchrun for N>1)
set -x
N=$1; W=$2; S=${3:-20}; shift; shift; shift; EXTRA="$@"
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 420 python bench.py --gpus 1 --no-cpu --workload $W --steps $S --warmup 3 $EXTRA > gpurun_out/scale_${W}_n$N.json 2>gpurun_out/scale_${W}_n$N.err
else
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps $S --warmup 3 $EXTRA > gpurun_out/scale_${W}_n$N.json 2>gpurun_out/scale_${W}_n$N.err
fi
echo rc=$?
python - <<PY
import json
for l in open('gpurun_out/scale_${W}_n$N.json'):
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print('RESULT', '$W', j['n_gpus'], 'users/s', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'spmm_ms', round(j['spmm']['ms'],3), 'score_ms', round(j['scoring']['ms'],3), 'e2e', j['e2e']['value'], 'mode', j['config']['parallelism'][:40], 'north_star', j.get('north_star_scale'))
PY
tail -3 gpurun_out/scale_${W}_n$N.err | cut -c1-300
