python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "v_cycle or reference_constant or multicolor_precond" > gpurun_out/c3_tests.log 2>&1; tail -3 gpurun_out/c3_tests.log
for ord in multicolor lexicographic; do
for fr in 1 0; do
  echo "== order $ord fused_restrict=$fr"
  MG_ORDER=$ord PA_MG_FUSED_RESTRICT=$fr timeout 300 python tools/mg_bench.py 512 4 2>&1 | grep "V-cycle\|MG-precond"
done; done > gpurun_out/c3_mg.log 2>&1
cat gpurun_out/c3_mg.log
