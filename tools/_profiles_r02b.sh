set -x
mkdir -p gpurun_out
python tools/train_ncu.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_train_step_launches_gpu_time.csv python tools/train_ncu.py > gpurun_out/ncu_train.log 2>&1
SPVD_BENCH_VERBOSE=1 python bench.py --steps 30 --warmup 3 2> gpurun_out/r02b_conv_table_per_layer.txt > gpurun_out/r02b_bench_line_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_line_reference.json 2>/dev/null
wc -l gpurun_out/*.csv; cut -c1-300 gpurun_out/r02b_bench_line_1gpu.json; cut -c1-300 gpurun_out/r02b_bench_line_reference.json
