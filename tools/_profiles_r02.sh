set -x
mkdir -p gpurun_out
python tools/profile_step.py 1 2 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_step_launches_gpu_time.csv python tools/profile_step.py 1 2 > gpurun_out/ncu_step.log 2>&1
python tools/train_ncu.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_launches_gpu_time.csv python tools/train_ncu.py > gpurun_out/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_conv_fwd_umma|k_conv_pairs|k_conv_sp" -c 80 -f -o gpurun_out/conv_r02 python tools/profile_step.py 1 2 > gpurun_out/ncu_conv.log 2>&1
ncu -i gpurun_out/conv_r02.ncu-rep --page raw --csv > gpurun_out/conv_r02_raw.csv 2>/dev/null
rm -f gpurun_out/conv_r02.ncu-rep   # (the merge back is capped at 64 MiB; the raw page is what profiles/ keeps)
python tools/hbm_kernels_bench.py 500 > gpurun_out/r02_hbm_kernels_gbs.txt 2>&1
SPVD_BENCH_VERBOSE=1 python bench.py --steps 30 --warmup 3 2> gpurun_out/r02_conv_table_per_layer.txt > gpurun_out/r02_bench_line_1gpu.json
du -sh gpurun_out; tail -3 gpurun_out/ncu_conv.log; wc -l gpurun_out/*.csv; cat gpurun_out/r02_bench_line_1gpu.json | cut -c1-400
