SWEEP_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_pat' -s 3 -c 1 -o gpurun_out/r02c_spmv7 python tools/spmv_sweep.py 7 512 > gpurun_out/r02c_ncu7.log 2>&1
SWEEP_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_pat' -s 3 -c 1 -o gpurun_out/r02c_spmv27 python tools/spmv_sweep.py 27 512 > gpurun_out/r02c_ncu27.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c_spmv7.ncu-rep > gpurun_out/r02c_ncu_spmv7_pat.txt 2>&1
python tools/ncu_summary.py gpurun_out/r02c_spmv27.ncu-rep > gpurun_out/r02c_ncu_spmv27_pat.txt 2>&1
head -24 gpurun_out/r02c_ncu_spmv7_pat.txt; head -12 gpurun_out/r02c_ncu_spmv27_pat.txt
rm -f gpurun_out/r02c_spmv7.ncu-rep gpurun_out/r02c_spmv27.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong --no-mg --no-hpcg27 > gpurun_out/r02c_launches.log 2>&1
