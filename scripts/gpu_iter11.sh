set -x
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6; do LGX_SPMM_VARIANT128=$v python scripts/spmm_sweep.py synth-10m; done 2>&1 | grep variant | tee gpurun_out/spmm_sweep128.jsonl
for c in 512 1024 2048; do LGX_CHUNK=$c python scripts/spmm_sweep.py synth-10m; done 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep128.jsonl
