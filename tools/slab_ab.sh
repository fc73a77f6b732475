#!/bin/bash
# A/B of the slab exchange variants on N GPUs (run under gpurun --gpus N): tools/slab_ab.sh N "<env1>" "<env2>" ...
N=$1; shift
port=29520
for envs in "$@"; do
  port=$((port+1))
  tag=$(echo "$envs" | tr ' =' '__')
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-ensemble-record ${SLAB_AB_FLAGS:---no-slab-check} ${SLAB_AB_EXTRA} > gpurun_out/slab_ab_${N}_${tag}.json 2> gpurun_out/slab_ab_${N}_${tag}.err
  echo "== $envs"; python tools/slabsum.py gpurun_out/slab_ab_${N}_${tag}.json
done
