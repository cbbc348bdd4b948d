"""Which option combination changes the emitted tokens at the headline shape (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
lib = sd._lib.lib()
V, g, B = 128256, 4, 256
gen = torch.Generator(device="cuda").manual_seed(1)
t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
tk = sd.sample_rows(d.reshape(B * g, V), None, seed=4321)[0].reshape(B, g)
def run(**opts):
    lib.specdec_set_option(b"reset", 1)
    for k, v in opts.items():
        assert lib.specdec_set_option(k.encode(), v) == 0
    r = sd.fused_verify(t, d, tk, None, None, seed=7, offset=3)
    torch.cuda.synchronize()
    lib.specdec_set_option(b"reset", 1)
    return r
ref = run(no_fused_tail=1, chunks=1)
for opts in [dict(), dict(tf_balance=0), dict(tail_slots=0), dict(tail_slots=0, tf_balance=0), dict(chunks=1), dict(chunks=1, tf_balance=0),
             dict(chunks=1, tail_slots=0), dict(tf_ch=21, tf_balance=0), dict(tf_ch=16, tf_balance=0)]:
    r = run(**opts)
    bad = (r.next_token != ref.next_token).nonzero().reshape(-1).tolist()
    badn = (r.n_accepted != ref.n_accepted).nonzero().reshape(-1).tolist()
    print(opts, "token mismatches:", bad[:10], len(bad), " n mismatches:", len(badn), flush=True)
    for b in bad[:3]:
        print("   seq", b, "n", int(ref.n_accepted[b]), "ref tok", int(ref.next_token[b]), "got", int(r.next_token[b]))
