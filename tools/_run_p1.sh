# launch list of one bench step (kernel shares), then ncu --set full of the new MG kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r02b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gs_color_tma|k_residual_restrict' -c 6 -o gpurun_out/r02b_mg python tools/prof_r02.py mg > gpurun_out/r02b_ncu_mg.log 2>&1
python tools/ncu_summary.py gpurun_out/r02b_mg.ncu-rep > gpurun_out/r02b_ncu_mg.txt 2>&1
tail -3 gpurun_out/r02b_ncu_mg.log; head -30 gpurun_out/r02b_ncu_mg.txt
rm -f gpurun_out/r02b_mg.ncu-rep
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 300 gpurun_out/r02b_bench_n1.err
