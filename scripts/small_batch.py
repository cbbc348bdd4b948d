"""A few verify steps at small batch (for ncu launch lists): B from $BS (default 1,32), gamma=4, V=128256, bf16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
V, g = 128256, 4
gen = torch.Generator(device="cuda").manual_seed(1)
for B in [int(x) for x in os.environ.get("BS", "1,32").split(",")]:
    t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
    d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
    tk = sd.sample_rows(d.reshape(B * g, V), None, seed=4321)[0].reshape(B, g)
    for i in range(int(os.environ.get("N", 4))):
        sd.fused_verify(t, d, tk, None, None, seed=7, offset=i)
    torch.cuda.synchronize()
print("done")
