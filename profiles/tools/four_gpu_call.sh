#!/bin/bash
# bench.py on four GPUs (torchrun, one rank per GPU): the sharded encode -> one stream -> sharded decode with the final build.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 10 --warmup 3 --quick > gpurun_out/b4_s2.json 2> gpurun_out/b4_s2.err
echo "rc=$?"; tail -c 1500 gpurun_out/b4_s2.json; tail -3 gpurun_out/b4_s2.err
