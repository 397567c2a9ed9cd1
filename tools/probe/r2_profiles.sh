set -x
python bench.py --torch-gpu-baseline --no-sub-records > gpurun_out/r2_bench_torch_gpu_baseline.json 2> gpurun_out/r2_bench_torch_gpu_baseline.err
python bench.py --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 420 --csv --log-file gpurun_out/r2_av_train_bf16_launches_final.csv python bench.py --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
python bench.py --workload video_resnet_lstm --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none -s 850 -c 320 --csv --log-file gpurun_out/r2_resnet_bf16_launches.csv python bench.py --workload video_resnet_lstm --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 60 -c 12 -f -o gpurun_out/r2_conv_igemm python bench.py --workload video_resnet_lstm --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu3.log 2>&1
