for st in 4 3 2; do
  LIPREAD_CONV_STAGES=$st python bench.py --workload video_resnet_lstm --no-cpu-baseline --no-sub-records --steps 10 --dump-ops gpurun_out/r2_ops_resnet_st$st.json > gpurun_out/r2_bench_resnet_st$st.json 2>/dev/null
done
