set -x
mkdir -p gpurun_out
for v in 0 20 21 22 23 24 25 26 27; do LGX_SPMM_VARIANT=$v python scripts/spmm_sweep.py amazon-book; done 2>&1 | grep variant | tee gpurun_out/spmm_sweep3.jsonl
for v in 0 21 22; do LGX_SPMM_VARIANT=$v python scripts/spmm_sweep.py gowalla; done 2>&1 | grep variant | tee -a gpurun_out/spmm_sweep3.jsonl
LGX_SPMM_VARIANT=21 timeout 600 python -m pytest tests/test_gpu_propagate.py -x -q 2>&1 | tail -3
