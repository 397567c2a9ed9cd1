export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_conv_igemm_gpu.py -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2_s2.log
timeout 600 python -m pytest tests/test_models_gpu.py -q -m gpu -k "resnet" 2>&1 | tail -5 >> gpurun_out/r2_s2.log
for f in 0 1; do LIPREAD_IGEMM_S2=$f timeout 300 python bench.py --workload video_resnet_lstm --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config 2 igemm_s2=$f', round(d['value']), round(d['ms_per_step'],3))" >> gpurun_out/r2_s2.log; done
