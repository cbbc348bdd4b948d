"""A few verify steps at small batch (for ncu launch lists / timing): B from $BS (default 1,32), gamma=4, V=128256, bf16."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
lib = sd._lib.lib()
V, g = 128256, 4
gen = torch.Generator(device="cuda").manual_seed(1)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("=")
    assert lib.specdec_set_option(k.encode(), int(v)) == 0
for B in [int(x) for x in os.environ.get("BS", "1,32").split(",")]:
    t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
    d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
    tk = sd.sample_rows(d.reshape(B * g, V), None, seed=4321)[0].reshape(B, g)
    N = int(os.environ.get("N", 4))
    for i in range(N):
        sd.fused_verify(t, d, tk, None, None, seed=7, offset=i)
    torch.cuda.synchronize()
    if os.environ.get("TIME"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(50):
            sd.fused_verify(t, d, tk, None, None, seed=7, offset=i)
        e1.record()
        host = (time.perf_counter() - t0) / 50 * 1e6
        torch.cuda.synchronize()
        gv = sd.GraphedVerify(t, d, tk, seed=7)
        for i in range(5):
            gv()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(50):
            gv()
        e3.record()
        torch.cuda.synchronize()
        print(f"B={B:3d} eager {e0.elapsed_time(e1) / 50 * 1e3:7.1f} us/step (host enqueue {host:6.1f} us)   graph replay {e2.elapsed_time(e3) / 50 * 1e3:7.1f} us/step", flush=True)
print("done")
