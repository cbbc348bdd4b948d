#!/bin/bash
# 1/2/4/8-GPU weak-scaling runs of the headline bench on one box (one rank per GPU, NCCL)
O=gpurun_out
nvidia-smi -L | wc -l
timeout 300 python bench.py --gpus 1 --no-sweep --no-cpu-baseline > $O/r1_scale_1gpu.json 2>/dev/null; echo rc=$?
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 200 --warmup 5 --no-sweep --no-cpu-baseline > $O/r1_scale_${N}gpu.json 2> $O/r1_scale_${N}gpu.err; echo N=$N rc=$?
done
for N in 1 2 4 8; do tail -1 $O/r1_scale_${N}gpu.json | cut -c1-330; done
