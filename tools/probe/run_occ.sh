run() { tag=$1; shift; for wl in video_resnet_lstm acv_late_fusion_mobile av_train; do env "$@" python bench.py --workload $wl --no-cpu-baseline --no-sub-records --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', '$wl', round(d['ms_per_step'],3))"; done; }
run base A=1
run wgrad1 LIPREAD_WGRAD_STAGES=1
run gemm2cta LIPREAD_GEMM_2CTA=1
run both LIPREAD_WGRAD_STAGES=1 LIPREAD_GEMM_2CTA=1
