#!/bin/bash
# usage: tools/run_dist.sh N script.py [args...]   -- one process per GPU on this node
N=$1; shift
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) "$@"
