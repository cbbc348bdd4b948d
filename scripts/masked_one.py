"""One top-k verify step at the headline shape (for ncu captures of rowsel_tma_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
V, g, B = 128256, 4, 256
gen = torch.Generator(device="cuda").manual_seed(1)
t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
m = dict(temperature=0.7, top_k=50, top_p=1.0, greedy=False)
tk = sd.sample_rows(d.reshape(B * g, V), None, seed=4321, **m)[0].reshape(B, g)
for i in range(2):
    sd.fused_verify(t, d, tk, None, None, seed=7, offset=i, **m)
torch.cuda.synchronize()
print("done")
