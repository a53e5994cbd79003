timeout 600 python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "strip or short_rows" > gpurun_out/s9_tests.log 2>&1; tail -3 gpurun_out/s9_tests.log
PA_GS_KERNEL=3 PA_GS_TRACE=1 MG_QUICK=1 timeout 300 python tools/mg_bench.py 128 1 2>&1 | grep "task 300 step 127\|task 0 step 127\|symmetric" | tail -5
PA_GS_KERNEL=3 MG_QUICK=1 timeout 300 python tools/mg_bench.py 512 4 2>&1 | grep "symmetric"
