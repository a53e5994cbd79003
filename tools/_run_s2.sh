PA_GS_KERNEL=3 PA_GS_TRACE=1 MG_QUICK=1 timeout 300 python tools/mg_bench.py 128 1 2>&1 | grep -v "^setup\|^order" | head -60 > gpurun_out/s2_trace.log
cat gpurun_out/s2_trace.log
