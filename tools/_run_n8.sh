( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err ) 2> gpurun_out/n8_bench.time; tail -3 gpurun_out/n8_bench.time
tail -c 400 gpurun_out/n8_bench.err
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1; numactl -H >> gpurun_out/n8_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/n8_topo.txt
