"""GPU probe of the megakernel: timing per batch size against the three-launch pipeline, result equality, option sweep."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specdec_b200 as sd
lib = sd._lib.lib()
V, g = 128256, 4
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(1)
Bmax = int(os.environ.get("BMAX", 256))
t = (3 * torch.randn(Bmax, g + 1, V, device=dev, generator=gen)).to(torch.bfloat16)
d = (t[:, :g].float() + 0.5 * torch.randn(Bmax, g, V, device=dev, generator=gen)).to(torch.bfloat16)
toks, _ = sd.sample_rows(d.reshape(Bmax * g, V), None, seed=4321)
toks = toks.reshape(Bmax, g)
t2 = t.clone(); d2 = d.clone()

def run(B, n=20, **opts):
    lib.specdec_set_option(b"reset", 1)
    for k, v in opts.items():
        assert lib.specdec_set_option(k.encode(), v) == 0, k
    sets = [(t, d), (t2, d2)]
    for i in range(3):
        r = sd.fused_verify(sets[i % 2][0][:B], sets[i % 2][1][:B], toks[:B], None, None, seed=7, offset=i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        r = sd.fused_verify(sets[i % 2][0][:B], sets[i % 2][1][:B], toks[:B], None, None, seed=7, offset=i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

for B in [int(x) for x in os.environ.get("BS", "1,8,32,64,128,256").split(",")]:
    if B > Bmax: continue
    ms0, r0 = run(B, n=int(os.environ.get("N", 20)), mega=0)
    ms1, r1 = run(B, n=int(os.environ.get("N", 20)), mega=1)
    same = all(torch.equal(a, b) for a, b in ((r0.n_accepted, r1.n_accepted), (r0.next_token, r1.next_token),
                                              (r0.accept_mask, r1.accept_mask), (r0.packed, r1.packed)))
    print(f"B={B:4d}  three-launch {ms0*1e3:8.1f} us   mega {ms1*1e3:10.1f} us   same={same}", flush=True)
def timeline(B):
    import numpy as np
    from specdec_b200 import ops
    lib.specdec_set_option(b"reset", 1)
    lib.specdec_set_option(b"mega_dbg", 1); lib.specdec_set_option(b"mega", 1)
    for i in range(3):
        sd.fused_verify(t[:B], d[:B], toks[:B], None, None, seed=7, offset=i)
    torch.cuda.synchronize()
    ws = list(ops._WS_CACHE.values())[0]
    n = 16 + 8 * B + 1024
    out = (ctypes.c_ulonglong * n)()
    assert lib.specdec_debug_timeline(ctypes.c_void_p(ws.data_ptr()), B, g, V, ctypes.cast(out, ctypes.c_void_p), n) == 0
    a = np.array(list(out), dtype=np.float64)
    t0 = a[0]
    glob = (a[1:4] - t0) / 1e3
    per = (a[16:16 + 8 * B].reshape(B, 8)[:, :8] - t0) / 1e3
    cta = np.array(list(out)[16 + 8 * B:16 + 8 * B + 592], dtype=np.uint64)
    isr = ((cta >> np.uint64(15)) & np.uint64(1)).astype(bool)
    smid = (cta & np.uint64(0x7FFF)).astype(int)
    order = np.argsort(~isr, kind="stable")
    smid, rend = smid[order], (cta >> np.uint64(16)).astype(np.float64)[order] / 1e3
    nr = int(isr.sum())
    import collections
    cnt = collections.Counter(smid[:nr].tolist())
    print("   R CTAs per SM histogram:", sorted(collections.Counter(cnt.values()).items()), " distinct SMs with R:", len(cnt),
          " smid of CTA 0..7:", smid[:8].tolist(), " CTA 148..151:", smid[148:152].tolist(), "CTA 296..299", smid[296:300].tolist())
    per_sm = {k: rend[:nr][smid[:nr] == k].max() for k in cnt}
    byc = collections.defaultdict(list)
    for k, c_ in cnt.items():
        byc[c_].append(per_sm[k])
    print("   R end time by #R CTAs on the SM:", {k: (round(min(v_), 1), round(float(np.median(v_)), 1), round(max(v_), 1)) for k, v_ in byc.items()})
    print(f"timeline B={B}: last R done {glob[0]:.1f} us, first R done {glob[2]:.1f}, kernel end {glob[1]:.1f} us")
    names = ["plan_start", "plan_pub", "item_first", "item_last", "norm_done", "fin_start", "fin_end", "A_done_last"]
    for k, nme in enumerate(names):
        col = per[:, k]
        col = col[(col > 0) & (col < 1e7)]
        if col.size:
            print(f"   {nme:10s} min {col.min():7.1f}  med {np.median(col):7.1f}  max {col.max():7.1f}   (n={col.size})")
    for b in list(range(0, B, max(1, B // 8)))[:8]:
        print("   seq", b, " ".join(f"{x:7.1f}" for x in per[b]))
    lib.specdec_set_option(b"reset", 1)

for B in [int(x) for x in os.environ.get("TL", "").split(",") if x]:
    timeline(B)
if os.environ.get("SWEEP"):
    B = Bmax
    for opts in [dict(mega_r=1), dict(mega_r=3), dict(mega_unit=2), dict(mega_unit=8), dict(mega_unit=16), dict(mega_keep_l2=0),
                 dict(mega_spc=12), dict(mega_spc=16), dict(mega_spc=20)]:
        ms1, r1 = run(B, mega=1, **opts)
        print(f"B={B} {opts}: mega {ms1*1e3:8.1f} us", flush=True)
lib.specdec_set_option(b"reset", 1)
