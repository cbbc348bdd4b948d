#!/bin/bash
# 1 -> 8 GPU scaling on one 8xB200 box: weak line (B=256 per GPU) with the strong-scaling pass (global B=256) inside.
O=gpurun_out
for N in 8 4 2; do
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 200 --warmup 5 --no-sweep --no-cpu-baseline --e2e-steps 3 > $O/r2_scale_${N}gpu.json 2> $O/r2_scale_${N}gpu.err
  echo "N=$N rc=$?"
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --steps 200 --warmup 5 --no-sweep --no-cpu-baseline --e2e-steps 2 --nccl-gather > $O/r2_scale_8gpu_nccl.json 2> $O/r2_scale_8gpu_nccl.err
echo "N=8 nccl rc=$?"
timeout -s KILL 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-sweep --no-cpu-baseline --e2e-steps 3 > $O/r2_scale_1gpu.json 2> $O/r2_scale_1gpu.err; echo "N=1 rc=$?"
python - <<'P'
import json
for n in ("1", "2", "4", "8", "8gpu_nccl"):
    f = f"gpurun_out/r2_scale_{n}gpu.json" if n.isdigit() else f"gpurun_out/r2_scale_{n}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("strong_scaling") or {}
        print(n, "weak %.2fM %.1fus" % (d["value"] / 1e6, d["ms_per_step"] * 1e3), "strong", st.get("value") and "%.2fM %.1fus" % (st["value"] / 1e6, st["ms_per_step"] * 1e3), d.get("gather", "")[:20])
    except Exception as e:
        print(n, "ERR", e)
P
