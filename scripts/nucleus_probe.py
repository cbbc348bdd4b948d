"""Timing probe: pure top-p verify on rows of different flatness (scale * randn logits)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import specdec_b200 as sd
B, g, V = 64, 4, 128256
gen = torch.Generator(device="cuda").manual_seed(0)
for scale in (3.0, 1.0, 0.3, 0.05):
    t = (scale * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
    d = (t[:, :g].float() + 0.3 * scale * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
    toks = torch.randint(V, (B, g), device="cuda", generator=gen)
    ua = torch.rand(B, g, device="cuda", generator=gen); us = torch.rand(B, device="cuda", generator=gen)
    for mode, kw in (("top_p0.9", dict(top_p=0.9)), ("plain", dict())):
        for _ in range(2):
            sd.fused_verify(t, d, toks, ua, us, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = sd.fused_verify(t, d, toks, ua, us, **kw)
        e1.record(); torch.cuda.synchronize()
        if mode != "plain":
            import ctypes
            st = (ctypes.c_ulonglong * 16)()
            sd._lib.lib().specdec_debug_stats(ctypes.cast(st, ctypes.c_void_p), 1)
            print("   hist stats", list(st)[:10], flush=True)
        print(f"scale={scale} {mode}: {e0.elapsed_time(e1)/5:.3f} ms/step (B={B}), mean accepted {float(r.n_accepted.float().mean()):.2f}", flush=True)
# bench-like rows: target 3*randn, drafter target + 0.5*randn, B=256
B = 256
t = (3.0 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
toks = torch.randint(V, (B, g), device="cuda", generator=gen)
ua = torch.rand(B, g, device="cuda", generator=gen); us = torch.rand(B, device="cuda", generator=gen)
r = sd.fused_verify(t, d, toks, ua, us, top_p=0.9); torch.cuda.synchronize()
import ctypes
st = (ctypes.c_ulonglong * 16)()
sd._lib.lib().specdec_debug_stats(ctypes.cast(st, ctypes.c_void_p), 1)
print("bench-like rows: hist stats", list(st)[:10], flush=True)
