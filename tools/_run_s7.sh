timeout 600 python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "strip or short_rows" > gpurun_out/s7_tests.log 2>&1; tail -3 gpurun_out/s7_tests.log
for v in 2 1; do
echo "== strip v$v"
PA_GS_KERNEL=3 PA_GS_STRIP_V=$v MG_QUICK=1 timeout 300 python tools/mg_bench.py 512 4 2>&1 | grep "symmetric"
done
