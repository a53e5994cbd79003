PA_GS_KERNEL=3 PA_GS_TRACE=1 MG_QUICK=1 timeout 300 python tools/mg_bench.py 128 1 2>&1 | grep "task 300 step 127\|task 0 step 127\|symmetric" | tail -5
