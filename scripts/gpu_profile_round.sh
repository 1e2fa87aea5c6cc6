# Round evidence: bench lines (ours bf16 / bf16x3, reference arm), ncu launch list, ncu --set full of the hot kernels.
#   scripts/gpu.sh --timeout 1500 -- 'bash scripts/gpu_profile_round.sh r2'
set -x
R=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/BENCH_${R}_ours.json 2> gpurun_out/bench_${R}.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/BENCH_${R}_reference.json 2>> gpurun_out/bench_${R}.err
python bench.py --mode bf16x3 --no-cpu --scale-leg off > gpurun_out/BENCH_${R}_ours_bf16x3.json 2>> gpurun_out/bench_${R}.err
python bench.py --steps 2 --warmup 1 --no-cpu --scale-leg off --min-seconds 0 --no-check > gpurun_out/plain_${R}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --scale-leg off --min-seconds 0 --no-check > gpurun_out/ncu_launch_${R}.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --scale-leg off --min-seconds 0 --no-check > gpurun_out/plain_${R}b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_spmm_fixed|k_score_topk_gq|k_rescore_topk|k_mask_buckets' -s 12 -c 6 -o gpurun_out/${R}_prof python bench.py --steps 2 --warmup 1 --no-cpu --scale-leg off --min-seconds 0 --no-check > gpurun_out/ncu_full_${R}.log 2>&1
tail -2 gpurun_out/ncu_full_${R}.log | cut -c1-200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_nvsmi.csv
