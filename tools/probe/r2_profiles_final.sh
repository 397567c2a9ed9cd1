export PYTHONPATH=$PWD
python bench.py --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1000 -c 400 --csv --log-file gpurun_out/r2_av_train_bf16_launches_final4.csv python bench.py --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
python bench.py --workload video_resnet_lstm --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/r2_resnet_bf16_launches_final4.csv python bench.py --workload video_resnet_lstm --no-sub-records --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1
