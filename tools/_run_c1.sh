python -m pytest tests/test_gpu_hpcg_mg.py -x -q -k "multicolor" > gpurun_out/c1_tests.log 2>&1; tail -5 gpurun_out/c1_tests.log
export MG_ORDER=multicolor
for cfg in "0 2 2 14" "1 2 2 14" "1 2 2 9" "1 2 2 27" "1 1 2 14" "1 4 2 14" "1 2 3 14" "1 2 4 14" "1 4 3 27"; do
  set -- $cfg
  echo "== color_kernel=$1 slices=$2 stages=$3 batch=$4"
  PA_GS_COLOR_KERNEL=$1 PA_GS_COLOR_SLICES=$2 PA_GS_COLOR_STAGES=$3 PA_GS_COLOR_BATCH=$4 MG_QUICK=1 timeout 300 python tools/mg_bench.py 512 4 2>&1 | grep -v "^setup\|^order"
done > gpurun_out/c1_sweep.log 2>&1
PA_GS_COLOR_KERNEL=1 timeout 300 python tools/mg_bench.py 512 4 > gpurun_out/c1_mg.log 2>&1
cat gpurun_out/c1_sweep.log; tail -3 gpurun_out/c1_mg.log
