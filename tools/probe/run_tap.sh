export PYTHONPATH=$PWD
python -m pytest tests/test_conv2d_gpu.py tests/test_bf16_kernels_gpu.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_tap.log
python -m pytest tests/test_models_gpu.py -q -m gpu -k "resnet" 2>&1 | tail -3 >> gpurun_out/r2_tap.log
python bench.py --workload video_resnet_lstm --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config 2', round(d['value']), round(d['ms_per_step'],3)); print(json.dumps(d['roofline'].get('time_by_op_ms', d['roofline']))[:1500])" >> gpurun_out/r2_tap.log
