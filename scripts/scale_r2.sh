#!/bin/bash
# 1 -> 8 GPU scaling on one 8xB200 box: weak line (B=256 per GPU) with the strong-scaling pass (global B=256) inside.
O=gpurun_out
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 200 --warmup 5 --no-sweep --no-cpu-baseline --e2e-steps 3 > $O/r2_scale_${N}gpu.json 2> $O/r2_scale_${N}gpu.err
  echo "N=$N rc=$?"; tail -c 900 $O/r2_scale_${N}gpu.json | head -c 900; echo
done
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-sweep --no-cpu-baseline --e2e-steps 3 > $O/r2_scale_1gpu.json 2> $O/r2_scale_1gpu.err; echo "N=1 rc=$?"
