set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 3000 gpurun_out/bench_bf16.json
python bench.py --mode bf16x3 --no-cpu > gpurun_out/bench_bf16x3.json 2>> gpurun_out/bench_bf16.err
python bench.py --mode fp32 --no-cpu --steps 5 > gpurun_out/bench_fp32.json 2>> gpurun_out/bench_bf16.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_bf16.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_spmm$|k_score_topk_tc|k_spmm_long' -s 8 -c 5 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
