"""A handful of small verify / helper calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
lib = sd._lib.lib()
gen = torch.Generator(device="cuda").manual_seed(1)
def case(B, g, V, dt):
    t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(dt)
    d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(dt)
    return t, d
modes = [dict(), dict(greedy=True), dict(temperature=0.7, top_k=50), dict(top_p=0.9), dict(temperature=0.7, top_k=50, top_p=0.9)]
for (B, g, V, dt) in [(5, 3, 32000, torch.bfloat16), (3, 4, 8192, torch.float32), (130, 2, 4096, torch.bfloat16)]:
    t, d = case(B, g, V, dt)
    for m in modes:
        tk = sd.sample_rows(d.reshape(B * g, V), None, seed=3, **m)[0].reshape(B, g)
        r = sd.fused_verify(t, d, tk, None, None, seed=7, offset=1, **m)
        if B <= 64 and not m.get("top_k") and not m.get("top_p"):
            for k, v in (("small_b", 64), ("tail_slots", 0)):
                lib.specdec_set_option(k.encode(), v)
                sd.fused_verify(t, d, tk, None, None, seed=7, offset=1, **m)
                lib.specdec_set_option(b"reset", 1)
    sd.fused_verify(t, None, tk, None, None, seed=7, greedy=True, flags=sd._lib.NGRAM)
    sd.ops.topk_ids(t[:, 0], 3)
torch.cuda.synchronize()
print("sanitize run done")
