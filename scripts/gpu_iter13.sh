set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_propagate.py tests/test_gpu_train.py -x -q 2>&1 | tail -3
for v in 0 21 24 26; do LGX_SPMM_VARIANT=$v python scripts/spmm_sweep.py amazon-book; done 2>&1 | grep variant | tee gpurun_out/spmm_sweep4.jsonl
python scripts/spmm_sweep.py gowalla 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep4.jsonl
python scripts/spmm_sweep.py synth-10m 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep4.jsonl
