export PYTHONPATH=$PWD
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_head.log
for f in 1; do LIPREAD_SMALL_LINEAR=$f python bench.py --steps 30 --warmup 5 --no-sub-records 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small_linear $f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'loss', d['final_loss'])" >> gpurun_out/r2_head.log; done
