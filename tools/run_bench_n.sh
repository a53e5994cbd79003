#!/bin/bash
# usage: tools/run_bench_n.sh N  -> gpurun_out/final_bench_nN.json (the driver's launch line for N > 1)
N=$1
if [ "$N" = "1" ]; then python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; fi
tail -c 300 gpurun_out/final_bench_n$N.err
