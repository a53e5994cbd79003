#!/bin/bash
# usage: tools/run_n.sh NGPUS PORT [bench args...]  -> prints a short summary of the bench JSON line
N=$1; PORT=$2; shift 2
OUT=$(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" 2>&1 | grep '^{"metric"' | tail -1)
if [ -z "$OUT" ]; then echo "NO JSON for $*"; exit 0; fi
echo "$OUT" >> gpurun_out/run_n_lines.jsonl
echo "$OUT" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('N=%d'%d['n_gpus'], ' '.join(sys.argv[1:]), '| value', round(d['value'],1), 'iters/s', round(d['cg_iters_per_sec'],2), 'spmv_ms', round(d['spmv_ms'],4), 'launches', d['gpu_launches'])" "$@"
