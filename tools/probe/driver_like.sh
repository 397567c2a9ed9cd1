s=$(date +%s); python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-s )) s" >> gpurun_out/r2_final_pytest.log
s=$(date +%s); python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s)-s )) s" >> gpurun_out/r2_final_smoke.log
s=$(date +%s); python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$? $(( $(date +%s)-s )) s" >> gpurun_out/r2_final_bench_reference.err
s=$(date +%s); python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$? $(( $(date +%s)-s )) s" >> gpurun_out/r2_final_bench.err
