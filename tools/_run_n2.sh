python -m pytest tests/test_multiprocess.py -x -q > gpurun_out/n2_tests.log 2>&1; tail -3 gpurun_out/n2_tests.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err ) 2> gpurun_out/n2_bench.time; tail -3 gpurun_out/n2_bench.time
tail -c 600 gpurun_out/n2_bench.err
