timeout 600 python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "strip" > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
for v in 0 1 2 3; do
  echo "== var $v trace 128^3"
  PA_GS_KERNEL=3 PA_GS_STRIP_VAR=$v PA_GS_TRACE=1 MG_QUICK=1 timeout 300 python tools/mg_bench.py 128 1 2>&1 | grep "task 300 step 127\|symmetric" | tail -3
  echo "== var $v 512^3"
  PA_GS_KERNEL=3 PA_GS_STRIP_VAR=$v MG_QUICK=1 timeout 300 python tools/mg_bench.py 512 2 2>&1 | grep "symmetric"
done > gpurun_out/s3_var.log 2>&1
cat gpurun_out/s3_var.log
