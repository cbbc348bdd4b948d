"""CPU: the canonical arithmetic (C oracle) against the RAW reference arithmetic (oracle/torch_port.py, torch fp32 in
the reference's op order) at the headline vocabulary V=128256: accept decisions and emitted tokens agree, and every
disagreement is a boundary case within the north_star tolerance (|u - p/q| <= 1e-5 relative, or a CDF shift <= 1e-5).
The full-size report (>= 10 k decisions) is scripts/ref_match_rate.py -> profiles/r2_ref_match_rate.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_match_rate_sample_and_every_mismatch_is_a_boundary_case(oracle_mod):
    import ref_match_rate
    for kind in ("randn", "llm"):
        st = ref_match_rate.run(24, "multinomial", kind, verbose=False)
        assert st["decisions"] == 96
        assert st["accept_unexplained"] == 0 and st["token_unexplained"] == 0, st
        assert st["accept_match_rate"] >= 0.97 and st["accepted_length_match_rate"] >= 0.9, st
        assert st["token_match_rate"] >= 0.8, st


def test_committed_report_has_no_unexplained_mismatch():
    p = os.path.join(ROOT, "profiles", "r2_ref_match_rate.json")
    rep = json.load(open(p))
    assert sum(r["decisions"] for r in rep) >= 10000
    for r in rep:
        assert r["accept_unexplained"] == 0 and r["token_unexplained"] == 0, r
