"""One megakernel launch at the headline shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
V, g, B = 128256, 4, int(os.environ.get("B", 256))
gen = torch.Generator(device="cuda").manual_seed(1)
t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
toks, _ = sd.sample_rows(d.reshape(B * g, V), None, seed=4321)
toks = toks.reshape(B, g)
for i in range(3):
    r = sd.fused_verify(t, d, toks, None, None, seed=7, offset=i)
torch.cuda.synchronize()
print("ok", int(r.n_accepted.sum()))
