"""GPU probe of the masked (top-k / top-p) modes: step time per mode / row shape with the candidate-gather route on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import specdec_b200 as sd
from cases import MODES
lib = sd._lib.lib()
V, g = 128256, 4
B = int(os.environ.get("B", 256))
N = int(os.environ.get("N", 10))
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(1)


def rows(kind):
    t = 3 * torch.randn(B, g + 1, V, device=dev, generator=gen)
    if kind == "peaked":
        t = t * 0.5
        idx = torch.randint(V, (B, g + 1, 24), device=dev, generator=gen)
        t.scatter_(2, idx, 12.0 + 8.0 * torch.rand(B, g + 1, 24, device=dev, generator=gen))
    if kind == "uniform":
        t = t * (0.05 / 3)
    d = (t[:, :g] + (0.5 if kind != "uniform" else 0.02) * torch.randn(B, g, V, device=dev, generator=gen)).to(torch.bfloat16)
    return t.to(torch.bfloat16), d


def timeit(t, d, mode, **opts):
    lib.specdec_set_option(b"reset", 1)
    for k, v in opts.items():
        assert lib.specdec_set_option(k.encode(), v) == 0
    m = MODES[mode]
    tk = sd.sample_rows(d.reshape(B * g, V), None, seed=4321, **m)[0].reshape(B, g)
    for i in range(3):
        sd.fused_verify(t, d, tk, None, None, seed=1, offset=i, **m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        r = sd.fused_verify(t, d, tk, None, None, seed=1, offset=i, **m)
    e1.record()
    torch.cuda.synchronize()
    lib.specdec_set_option(b"reset", 1)
    return e0.elapsed_time(e1) / N, r


for kind in os.environ.get("KINDS", "randn,peaked,uniform").split(","):
    t, d = rows(kind)
    for mode in os.environ.get("MODES", "topk50,topk50_p0.9,nucleus0.9").split(","):
        ms1, r1 = timeit(t, d, mode)
        ms0, r0 = timeit(t, d, mode, no_rowsel=1, no_klist=1)
        if os.environ.get("PROBE"):
            msp, _ = timeit(t, d, mode, rowsel_probe=1)
            print(f"   (selector idle: {msp*1e3:8.1f} us incl. the fallback kernels doing every row)")
        same = torch.equal(r0.n_accepted, r1.n_accepted) and torch.equal(r0.next_token, r1.next_token)
        print(f"B={B} {kind:8s} {mode:12s} rowsel+lists {ms1*1e3:8.1f} us   round-1 kernels {ms0*1e3:8.1f} us   same={same}", flush=True)
