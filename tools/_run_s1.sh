timeout 600 python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "strip or (short_rows and 3)" > gpurun_out/s1_tests.log 2>&1; tail -5 gpurun_out/s1_tests.log
for k in 3 0; do
  echo "== gs_kernel=$k"
  PA_GS_KERNEL=$k MG_QUICK=1 timeout 300 python tools/mg_bench.py 512 4 2>&1 | grep -v "^setup\|^order"
done > gpurun_out/s1_sweep.log 2>&1
cat gpurun_out/s1_sweep.log
