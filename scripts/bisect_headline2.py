"""The failing oracle comparison at the headline shape: which sequences differ, and under which options (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import specdec_b200 as sd
from oracle import oracle
oracle.build()
from test_verify_gpu import _full_size_case
from cases import MODES
lib = sd._lib.lib()
case = _full_size_case(256, 4, 128256, "multinomial", 1, oracle)
m = MODES["multinomial"]
o = oracle.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
def run(**opts):
    lib.specdec_set_option(b"reset", 1)
    for k, v in opts.items():
        assert lib.specdec_set_option(k.encode(), v) == 0
    r = sd.fused_verify(*args, **m)
    torch.cuda.synchronize()
    lib.specdec_set_option(b"reset", 1)
    return r
for opts in [dict(), dict(tf_balance=0), dict(tail_slots=0), dict(tail_slots=0, tf_balance=0), dict(chunks=1), dict(no_fused_tail=1), dict(no_fused_tail=1, chunks=1)]:
    r = run(**opts)
    nt = r.next_token.cpu().numpy()
    bad = np.nonzero(nt != o.next_token)[0].tolist()
    badn = np.nonzero(r.n_accepted.cpu().numpy() != o.n_accepted)[0].tolist()
    print(opts, "token mismatches:", bad[:10], len(bad), " n mismatches:", len(badn), flush=True)
    for b in bad[:4]:
        print("   seq", b, "n", int(o.n_accepted[b]), "oracle tok", int(o.next_token[b]), "got", int(nt[b]), "u_sample", float(case["u_sample"][b]))
