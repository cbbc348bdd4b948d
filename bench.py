#!/usr/bin/env python
"""bench.py -- verified draft tokens/s of the fused speculative-sampling verify path.

One "step" = one verify step (specdec::verify through the torch op / C ABI) over one batch of
synthetic logits of BASELINE.json's shape: B=256 sequences, gamma=4 drafts, V=128256, bf16
(configs[1]).  `value` is timed with inputs resident in HBM; `e2e` is the same call fed from pinned
HOST buffers with the H2D copy of the logits and the D2H read of the packed result inside the timed
region.  N>1: one process per GPU (torchrun), sequences sharded by rank (weak scaling: B per rank
fixed), the only collective is the all-gather of the packed int32 results.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--mode multinomial|greedy|topk50|nucleus0.9|...] [--dtype bf16|f32|f16] [--sweep]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODES = {
    "greedy": dict(temperature=1.0, top_k=0, top_p=1.0, greedy=True),
    "multinomial": dict(temperature=1.0, top_k=0, top_p=1.0, greedy=False),
    "temp0.7": dict(temperature=0.7, top_k=0, top_p=1.0, greedy=False),
    "topk50": dict(temperature=0.7, top_k=50, top_p=1.0, greedy=False),
    "nucleus0.9": dict(temperature=1.0, top_k=0, top_p=0.9, greedy=False),
    "topk50_p0.9": dict(temperature=0.7, top_k=50, top_p=0.9, greedy=False),
}
ESIZE = {"bf16": 2, "f16": 2, "f32": 4}


def alg_bytes(B, gamma, V, dtype):
    """every target/drafter logit read exactly once, nothing V-sized written."""
    return B * V * (gamma * 2 * ESIZE[dtype] + ESIZE[dtype])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, 2 ms period, own thread)."""

    def __init__(self, index=0):
        self.index, self.rows, self.stop_flag, self.t, self.err = index, [], False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = f"nvml unavailable: {e}"
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(mhz), int(rs)))
            except Exception as e:  # pragma: no cover
                self.err = str(e)
                return
            time.sleep(0.002)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not started"]}
        self.stop_flag = True
        self.t.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({k for k, bit in names.items() for r in self.rows if r[1] & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def make_inputs(torch, B, gamma, V, dtype, sigma, seed, device, nbuf):
    """nbuf independent input sets generated on the device (synthetic, seeded)."""
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dtype]
    sets = []
    for i in range(nbuf):
        g = torch.Generator(device=device).manual_seed(1234 + seed * 131 + i)
        t = torch.empty(B, gamma + 1, V, dtype=dt, device=device)
        d = torch.empty(B, gamma, V, dtype=dt, device=device)
        for b0 in range(0, B, 32):  # chunked: keeps the fp32 temporaries small
            b1 = min(B, b0 + 32)
            tf = 3.0 * torch.randn(b1 - b0, gamma + 1, V, device=device, generator=g)
            t[b0:b1] = tf.to(dt)
            d[b0:b1] = (tf[:, :gamma] + sigma * torch.randn(b1 - b0, gamma, V, device=device, generator=g)).to(dt)
            del tf
        sets.append((t, d))
    return sets


def config_dict(args, world):
    """Same dict for both arms (the driver compares them)."""
    B, g, V, dtype = args.B, args.gamma, args.V, args.dtype
    per = "B=%d/gpu" % B if args.scaling == "weak" else "B=%d global (%d/gpu)" % (B, B // world)
    ab = alg_bytes(B if args.scaling == "weak" else B // world, g, V, dtype)
    return {"workload": f"synthetic logits verify: {per} gamma={g} V={V} {dtype} mode={args.mode} "
                        f"sigma={args.sigma} (BASELINE.json configs[1])",
            "l2": f"inputs {ab / 1e6:.0f} MB per step and GPU > 126 MB L2, rotated over {args.nbuf} buffers"
                  if ab > 126e6 * 1.5 else
                  f"inputs {ab / 1e6:.0f} MB per step and GPU, rotated over {args.nbuf} buffers (sum > 126 MB L2)",
            "parallelism": f"dp{world} (sequences sharded by rank, all-gather of packed results)"}


def count_our_launches(torch, step_fn, n=2):
    """Kernels of the library (namespace specdec::) observed by CUPTI (torch.profiler) over n steps -> (per step, by name)."""
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            step_fn(i)
        torch.cuda.synchronize()
    by = {}
    for ev in prof.key_averages():
        if "specdec::" in ev.key:
            short = ev.key.split("specdec::")[1].split("<")[0].split("(")[0]
            by[short] = by.get(short, 0) + ev.count
    tot = sum(by.values())
    if tot == 0:
        raise RuntimeError("profiler saw no specdec:: kernels")
    return tot / n, {k: v / n for k, v in by.items()}


def numa_local(local_gpu):
    """Pin this process to the CPUs of the GPU's NUMA node before the pinned host buffers are allocated (first touch):
    with 8 ranks copying 591 MB per step each, buffers on one node halve the H2D rate of the far GPUs."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_gpu).pci_bus_id
        dom = torch.cuda.get_device_properties(local_gpu).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_gpu).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cl = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import specdec_b200 as sd
    from specdec_b200 import _lib as L

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Bglob = args.B
    if args.scaling == "strong":
        assert args.B % world == 0, "strong scaling shards the global batch evenly"
        args_B_local = args.B // world
    else:
        args_B_local = args.B
    B, g, V, dtype = args_B_local, args.gamma, args.V, args.dtype
    mode = MODES[args.mode]
    nbuf = args.nbuf
    sets = make_inputs(torch, B, g, V, dtype, args.sigma, rank, dev, nbuf)
    # draft tokens from the drafter's own distribution (fused processor+sample op), as in real use
    toks = []
    for (t, d) in sets:
        tk, _ = sd.sample_rows(d.reshape(B * g, V), None, seed=4321, offset=0, seq_id0=rank * B * g, **mode)
        toks.append(tk.reshape(B, g))
    seq0 = rank * B

    pending = []
    # the all-gather of the packed results: peer-to-peer stores by a kernel of ours over NVLink (dist.PeerGather) when
    # symmetric memory is available, else one NCCL all-gather per step on NCCL's stream
    pgather, gather_kind = None, "none (single GPU)"
    if world > 1:
        gather_kind = "nccl all_gather_into_tensor, async"
        if not args.nccl_gather:
            try:
                pgather = sd.dist.PeerGather(world * B, g + 2, slots=8, overlap=True)
                gather_kind = ("p2p stores over NVLink by specdec_peer_publish (symmetric memory) on a side stream, "
                               "reader lags 2 steps")
            except Exception as ex:
                print(f"[bench] PeerGather unavailable ({str(ex)[:120]}): NCCL all-gather", file=sys.stderr)
    pg_step = [0]

    def step(i, ev=None):
        t, d = sets[i % nbuf]
        r = sd.fused_verify(t, d, toks[i % nbuf], None, None, seed=2025, offset=i, seq_id0=seq0, **mode)
        if pgather is not None:
            k = pg_step[0]
            out = pgather.publish(r.packed, k, k - 2)  # + every rank's results of two steps ago, in stream order
            pg_step[0] = k + 1
            return out
        if world > 1:
            # the gathered result is global bookkeeping; a rank continues on its own sequences, so the
            # 768-byte all-gather runs on NCCL's stream and overlaps the next verify step
            out, work = sd.dist.all_gather_packed(r.packed, world * B, async_op=True)
            pending.append(work)
            if len(pending) > 32:  # (bounds the outstanding collectives; the verify stream never waits for the newest)
                pending.pop(0).wait()
            return out
        return r.packed

    def drain():
        while pending:
            pending.pop(0).wait()
        if pgather is not None:
            pgather.sync_reader()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    sync_all()

    # ---- timed region: EXACTLY K plain steps, device-resident inputs, nothing else on the stream
    lib = L.lib()
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync_all()
    e0.record()
    for i in range(K):
        step(args.warmup + i)
    drain()
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt[0])
    ms_step = ms_total / K
    value = world * B * g / (ms_step * 1e-3)

    # ---- per-kernel split, instrumented passes OUTSIDE the timed region (the C-ABI hook records CUDA events on the
    # verify stream before the row kernel(s), after the last row-kernel launch and at the end of the call; six
    # event records per step cost ~10 us, which is why the K timed steps above run without them)
    def instrumented(n):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        for ev3 in evs:
            for e in ev3:
                e.record()  # creates the underlying cudaEvent_t
        sync_all()
        for i in range(n):
            lib.specdec_set_profile_events(evs[i][0].cuda_event, evs[i][1].cuda_event, evs[i][2].cuda_event)
            step(i)
        lib.specdec_set_profile_events(None, None, None)
        drain()
        sync_all()
        return (sum(evs[i][0].elapsed_time(evs[i][1]) for i in range(n)) / n,
                sum(evs[i][1].elapsed_time(evs[i][2]) for i in range(n)) / n)

    K2 = min(K, 50)
    t_rowstats, t_decide = instrumented(K2)
    # bf16/fp16 batches run as two chunks on two streams (row kernel of chunk 1 overlaps the exact tail of chunk 0),
    # so the span above covers two row-kernel launches plus the overlapped tail.  The roofline figure wants the
    # dominant kernel by itself: one more pass with the chunk pipelining off, where one row-kernel launch reads all
    # the algorithmic bytes and nothing else runs beside it.
    t_rowstats_in_step, t_decide_in_step = t_rowstats, t_decide
    pipelined = (dtype != "f32" and B >= 128 and mode["top_k"] == 0 and mode["top_p"] >= 1.0)
    if pipelined:
        lib.specdec_set_option(b"chunks", 1)
        try:
            for i in range(3):
                step(i)
            drain()
            t_rowstats, t_decide = instrumented(K2)
        finally:
            lib.specdec_set_option(b"chunks", 0)  # back to the library default

    # ---- the same step replayed from a CUDA graph (the C-ABI call is capturable, fork/join of the library's
    # auxiliary stream included): what a decode loop that graphs its step would see.  Reported, not the headline.
    ms_graph = None
    if world == 1:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(2):
                    sd.fused_verify(sets[0][0], sets[0][1], toks[0], None, None, seed=2025, offset=i, **mode)
            torch.cuda.current_stream().wait_stream(side)
            graphs = []
            for j in range(nbuf):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    rg = sd.fused_verify(sets[j][0], sets[j][1], toks[j], None, None, seed=2025, offset=j, **mode)
                graphs.append((gph, rg))
            for i in range(6):
                graphs[i % nbuf][0].replay()
            torch.cuda.synchronize()
            e0.record()
            for i in range(K):
                graphs[i % nbuf][0].replay()
            e1.record()
            torch.cuda.synchronize()
            ms_graph = e0.elapsed_time(e1) / K
            del graphs
        except Exception as ex:  # reported as absent, never fatal for the bench line
            print(f"[bench] graph replay skipped: {ex}", file=sys.stderr)

    # ---- kernels of ours per step, OBSERVED (CUPTI through torch.profiler, two untimed steps)
    launches_src = "observed: CUPTI kernel records named specdec::* over 2 untimed steps x K"
    try:
        per_step, by_name = count_our_launches(torch, step)
        drain()
    except Exception as ex:
        per_step, by_name = 3 * (2 if pipelined else 1), {}
        launches_src = f"computed (profiler unavailable: {str(ex)[:80]})"
    sync_all()

    # ---- two independent batches in flight on two streams (a serving engine that pipelines micro-batches): the exact
    # tail of one batch overlaps the HBM-bound row pass of the other.  Reported, NOT the headline: per-step latency
    # is unchanged, only the throughput of back-to-back independent steps rises.
    ms_two = None
    if world == 1:
        try:
            s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
            for s_ in s2:
                s_.wait_stream(torch.cuda.current_stream())
            def two(n):
                for i in range(n):
                    with torch.cuda.stream(s2[i & 1]):
                        sd.fused_verify(sets[i % nbuf][0], sets[i % nbuf][1], toks[i % nbuf], None, None, seed=2025,
                                        offset=i, seq_id0=seq0, **mode)
            two(4)
            torch.cuda.synchronize()
            e0.record()
            for s_ in s2:
                s_.wait_event(e0)
            two(K)
            for s_ in s2:
                torch.cuda.current_stream().wait_stream(s_)
            e1.record()
            torch.cuda.synchronize()
            ms_two = e0.elapsed_time(e1) / K
        except Exception as ex:
            print(f"[bench] two-batch pipelining skipped: {ex}", file=sys.stderr)

    # ---- strong scaling (BASELINE.json configs[4]: the GLOBAL batch sharded across the GPUs) next to the weak line
    strong = None
    if world > 1 and args.scaling == "weak" and Bglob % world == 0:
        Bs_ = Bglob // world
        spg = None
        if pgather is not None:
            try:
                spg = sd.dist.PeerGather(world * Bs_, g + 2, slots=8, overlap=True)
            except Exception:
                spg = None
        sk = [0]
        # B/world sequences per GPU is the small-batch regime where the eager call is bound by its ~50 us of host work:
        # the strong-scaling pass replays the step from a CUDA graph (sd.GraphedVerify, device-resident Philox offset)
        gvs = None
        try:
            if Bs_ > 64:
                raise RuntimeError("eager step is faster above 64 sequences per GPU")
            gvs = [sd.GraphedVerify(sets[j][0][:Bs_], sets[j][1][:Bs_], toks[j][:Bs_], seed=2025, seq_id0=rank * Bs_, **mode)
                   for j in range(nbuf)]
        except Exception as ex:
            print(f"[bench] strong pass: eager calls ({str(ex)[:100]})", file=sys.stderr)

        def sstep(i):
            if gvs is not None:
                r = gvs[i % nbuf]()
            else:
                t, d = sets[i % nbuf]
                r = sd.fused_verify(t[:Bs_], d[:Bs_], toks[i % nbuf][:Bs_], None, None, seed=2025, offset=i,
                                    seq_id0=rank * Bs_, **mode)
            if spg is not None:
                spg.publish(r.packed, sk[0], sk[0] - 2)
                sk[0] += 1
                return
            out_, work = sd.dist.all_gather_packed(r.packed, world * Bs_, async_op=True)
            pending.append(work)
            if len(pending) > 32:
                pending.pop(0).wait()
        def sdrain():
            drain()
            if spg is not None:
                spg.sync_reader()
        for i in range(5):
            sstep(i)
        sdrain()
        sync_all()
        e0.record()
        for i in range(K):
            sstep(i)
        sdrain()
        e1.record()
        sync_all()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_s = float(tt[0]) / K
        strong = {"global_B": Bglob, "B_per_gpu": Bs_, "ms_per_step": ms_s, "value": Bglob * g / (ms_s * 1e-3),
                  "unit": "tokens/s", "step": "CUDA-graph replay (GraphedVerify)" if gvs is not None else "eager",
                  "note": "same timing rules; value(N) / value at N=1 of the weak line = strong-scaling speed-up"}

    # ---- e2e: HOST logits (pinned) -> H2D -> verify -> D2H packed result, all inside the timed region
    t0, d0 = sets[0]
    numa = numa_local(local) if world > 1 else None
    ht = torch.empty(t0.shape, dtype=t0.dtype, pin_memory=True).copy_(t0)
    hd = torch.empty(d0.shape, dtype=d0.dtype, pin_memory=True).copy_(d0)
    htok = torch.empty(toks[0].shape, dtype=toks[0].dtype, pin_memory=True).copy_(toks[0])
    hout = torch.empty((world * B if world > 1 else B, g + 2), dtype=torch.int32, pin_memory=True)
    dt_, dd_, dk_ = torch.empty_like(t0), torch.empty_like(d0), torch.empty_like(toks[0])

    def e2e_step(i):
        dt_.copy_(ht, non_blocking=True)
        dd_.copy_(hd, non_blocking=True)
        dk_.copy_(htok, non_blocking=True)
        r = sd.fused_verify(dt_, dd_, dk_, None, None, seed=2025, offset=i, seq_id0=seq0, **mode)
        pk = sd.dist.all_gather_packed(r.packed, world * B) if world > 1 else r.packed
        hout.copy_(pk, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the result on the host

    Ke = max(2, min(K, args.e2e_steps))
    for i in range(2):
        e2e_step(i)
    sync_all()
    e0.record()
    for i in range(Ke):
        e2e_step(i)
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt[0])
    e2e_val = world * B * g / (ms_e2e / Ke * 1e-3)
    h2d = ht.numel() * ht.element_size() + hd.numel() * hd.element_size() + htok.numel() * 8
    d2h = hout.numel() * 4

    # ---- secondary sweep (not the headline): other processor modes on the same inputs, few steps each
    sweep = {}
    if world == 1 and not args.no_sweep:
        for mname in ("greedy", "topk50", "topk50_p0.9", "nucleus0.9"):
            md = MODES[mname]
            tk = [sd.sample_rows(d.reshape(B * g, V), None, seed=4321, offset=0, seq_id0=0, **md)[0].reshape(B, g)
                  for (t, d) in sets]
            for i in range(3):
                sd.fused_verify(sets[i % nbuf][0], sets[i % nbuf][1], tk[i % nbuf], None, None, seed=1, offset=i, **md)
            torch.cuda.synchronize()
            ks = 5 if mname == "nucleus0.9" else 20
            e0.record()
            for i in range(ks):
                sd.fused_verify(sets[i % nbuf][0], sets[i % nbuf][1], tk[i % nbuf], None, None, seed=1, offset=i, **md)
            e1.record()
            torch.cuda.synchronize()
            msm = e0.elapsed_time(e1) / ks
            sweep[mname] = {"tokens_per_s": B * g / (msm * 1e-3), "ms_per_step": msm,
                            "step_frac_of_hbm_peak": alg_bytes(B, g, V, dtype) / (msm * 1e-3) / 1e9 / peaks()[0]}

        # LLM-like (peaked) rows: a few dominant tokens, so the nucleus is small (the flat 3*randn rows above are
        # the worst case for top-p: ~5 % of the vocabulary is kept)
        gpk = torch.Generator(device=dev).manual_seed(99)
        tp_ = sets[0][0].float() * 0.5
        idx = torch.randint(V, (B, g + 1, 24), device=dev, generator=gpk)
        tp_.scatter_(2, idx, 12.0 + 8.0 * torch.rand(B, g + 1, 24, device=dev, generator=gpk))
        dp_ = (tp_[:, :g] + 0.3 * torch.randn(B, g, V, device=dev, generator=gpk)).to(sets[0][0].dtype)
        tp_ = tp_.to(sets[0][0].dtype)
        for mname in ("nucleus0.9", "topk50_p0.9"):
            md = MODES[mname]
            tk = sd.sample_rows(dp_.reshape(B * g, V), None, seed=4321, offset=0, seq_id0=0, **md)[0].reshape(B, g)
            for i in range(3):
                sd.fused_verify(tp_, dp_, tk, None, None, seed=1, offset=i, **md)
            torch.cuda.synchronize()
            e0.record()
            for i in range(10):
                sd.fused_verify(tp_, dp_, tk, None, None, seed=1, offset=i, **md)
            e1.record()
            torch.cuda.synchronize()
            msm = e0.elapsed_time(e1) / 10
            sweep[mname + "_peaked_rows"] = {"tokens_per_s": B * g / (msm * 1e-3), "ms_per_step": msm,
                                             "step_frac_of_hbm_peak": alg_bytes(B, g, V, dtype) / (msm * 1e-3) / 1e9 / peaks()[0]}
        del tp_, dp_
        # near-uniform rows (what random-init models emit, BASELINE.json configs[2]): top-p 0.9 keeps ~85 % of the
        # vocabulary; the cost of the histogram select does not depend on the size of the nucleus
        gun = torch.Generator(device=dev).manual_seed(98)
        tu_ = (0.05 * torch.randn(B, g + 1, V, device=dev, generator=gun))
        du_ = (tu_[:, :g] + 0.02 * torch.randn(B, g, V, device=dev, generator=gun)).to(sets[0][0].dtype)
        tu_ = tu_.to(sets[0][0].dtype)
        md = MODES["nucleus0.9"]
        tk = sd.sample_rows(du_.reshape(B * g, V), None, seed=4321, offset=0, seq_id0=0, **md)[0].reshape(B, g)
        for i in range(3):
            sd.fused_verify(tu_, du_, tk, None, None, seed=1, offset=i, **md)
        torch.cuda.synchronize()
        e0.record()
        for i in range(5):
            sd.fused_verify(tu_, du_, tk, None, None, seed=1, offset=i, **md)
        e1.record()
        torch.cuda.synchronize()
        msm = e0.elapsed_time(e1) / 5
        sweep["nucleus0.9_near_uniform_rows"] = {"tokens_per_s": B * g / (msm * 1e-3), "ms_per_step": msm,
                                                "step_frac_of_hbm_peak": alg_bytes(B, g, V, dtype) / (msm * 1e-3) / 1e9 / peaks()[0]}
        del tu_, du_
        # latency at small batch (headline mode)
        for Bs in (1, 8, 32, 64):
            t, d = sets[0]
            for i in range(3):
                sd.fused_verify(t[:Bs], d[:Bs], toks[0][:Bs], None, None, seed=1, offset=i, **mode)
            torch.cuda.synchronize()
            e0.record()
            for i in range(20):
                sd.fused_verify(t[:Bs], d[:Bs], toks[0][:Bs], None, None, seed=1, offset=i, **mode)
            e1.record()
            torch.cuda.synchronize()
            msm = e0.elapsed_time(e1) / 20
            sweep[f"{args.mode}_B{Bs}"] = {"tokens_per_s": Bs * g / (msm * 1e-3), "ms_per_step": msm,
                                           "step_frac_of_hbm_peak": alg_bytes(Bs, g, V, dtype) / (msm * 1e-3) / 1e9 / peaks()[0]}
            # the same step through GraphedVerify (captured once, device-resident Philox offset bumped inside the
            # graph): one cudaGraphLaunch per step instead of four enqueues + Python argument handling
            try:
                gv = [sd.GraphedVerify(sets[j][0][:Bs], sets[j][1][:Bs], toks[j][:Bs], seed=1, **mode) for j in range(nbuf)]
                for i in range(6):
                    gv[i % nbuf]()
                torch.cuda.synchronize()
                e0.record()
                for i in range(50):
                    gv[i % nbuf]()
                e1.record()
                torch.cuda.synchronize()
                msg = e0.elapsed_time(e1) / 50
                sweep[f"{args.mode}_B{Bs}"]["graph_replay_ms_per_step"] = msg
                del gv
            except Exception as ex:
                sweep[f"{args.mode}_B{Bs}"]["graph_replay_error"] = str(ex)[:120]

        # ---- the rest of BASELINE.json configs[1] (dtype / vocabulary / gamma axes) and the configs[2] shape
        # (B=64, gamma=4, nucleus p=0.9), each on its own inputs (two rotated sets; sets below the L2 size say so)
        def sweep_case(name, Bc, gc, Vc, dtc, mname, steps_=20):
            md = MODES[mname]
            ss = make_inputs(torch, Bc, gc, Vc, dtc, args.sigma, 7, dev, 2)
            tk = [sd.sample_rows(d_.reshape(Bc * gc, Vc), None, seed=4321, offset=0, seq_id0=0, **md)[0].reshape(Bc, gc)
                  for (_, d_) in ss]
            for i in range(3):
                sd.fused_verify(ss[i % 2][0], ss[i % 2][1], tk[i % 2], None, None, seed=1, offset=i, **md)
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps_):
                sd.fused_verify(ss[i % 2][0], ss[i % 2][1], tk[i % 2], None, None, seed=1, offset=i, **md)
            e1.record()
            torch.cuda.synchronize()
            msm = e0.elapsed_time(e1) / steps_
            abc = alg_bytes(Bc, gc, Vc, dtc)
            sweep[name] = {"tokens_per_s": Bc * gc / (msm * 1e-3), "ms_per_step": msm,
                           "step_frac_of_hbm_peak": abc / (msm * 1e-3) / 1e9 / peaks()[0],
                           "inputs_mb": abc / 1e6, "l2": "2 rotated sets" + ("" if 2 * abc > 126e6 else " (fit the 126 MB L2)")}
            del ss
        try:
            sweep_case("configs2_B64_g4_V128256_bf16_nucleus0.9", 64, 4, 128256, "bf16", "nucleus0.9", 10)
            sweep_case("B256_g4_V128256_f32_multinomial", 256, 4, 128256, "f32", "multinomial")
            sweep_case("B256_g4_V32000_bf16_multinomial", 256, 4, 32000, "bf16", "multinomial")
            sweep_case("B256_g1_V128256_bf16_multinomial", 256, 1, 128256, "bf16", "multinomial")
            sweep_case("B256_g8_V128256_bf16_multinomial", 256, 8, 128256, "bf16", "multinomial")
            sweep_case("B256_g4_V32000_f32_topk50", 256, 4, 32000, "f32", "topk50")
        except Exception as ex:
            sweep["sweep_case_error"] = str(ex)[:200]

        # ---- the two companion kernels of the path, reported as secondary numbers
        try:
            # KV rollback: Llama-3-8B-like static cache slice, 8 layers x (K,V), B=64, H_kv=8, S=2048, D=128, bf16
            Lk, Bk, Hk, Sk, Dk = 8, 64, 8, 2048, 128
            kv = [torch.zeros(Bk, Hk, Sk, Dk, dtype=torch.bfloat16, device=dev) for _ in range(2 * Lk)]
            lens = torch.full((Bk,), Sk, dtype=torch.int32, device=dev)
            disc = torch.randint(1, g + 2, (Bk,), device=dev, dtype=torch.int32)
            cache = sd.StaticKVCache(kv, lens)
            reps = 20
            for zf in (False, True):
                cache.rollback(disc, zero_fill=zf)
                torch.cuda.synchronize()
                e0.record()
                for i in range(reps):
                    cache.seq_lens.fill_(Sk)
                    cache.rollback(disc, zero_fill=zf)
                e1.record()
                torch.cuda.synchronize()
                msk = e0.elapsed_time(e1) / reps
                zbytes = int(disc.sum()) * Hk * Dk * 2 * 2 * Lk
                sweep["prune_kv" + ("_zero_fill" if zf else "")] = {
                    "ms_per_call": msk, "bytes_zeroed": zbytes if zf else 0,
                    "note": f"{2*Lk} tensors [64,8,2048,128] bf16, per-sequence discard 1..{g+1} (incl. the lens.fill_ launch)"}
            del kv, cache
            # n-gram tables: per-sequence tables, B=128 sequences, prompt 256 tokens, gamma=6 chained lookups
            Bn, Ln = 128, 256
            st = sd.NGramStorage(4, V, n_tables=Bn, grams_per_table=4096, counts_per_table=8192, device=dev)
            ids = torch.randint(0, 50, (Bn, Ln), device=dev)
            tabs = torch.arange(Bn, dtype=torch.int32, device=dev)
            st.initialize(ids, table_ids=tabs)
            fb = torch.zeros(Bn, 6, dtype=torch.long, device=dev)
            st.lookup_chain(ids, 6, table_ids=tabs, fallback=fb)
            torch.cuda.synchronize()
            e0.record()
            for i in range(reps):
                st.lookup_chain(ids, 6, table_ids=tabs, fallback=fb)
            e1.record()
            torch.cuda.synchronize()
            sweep["ngram_lookup_chain"] = {"ms_per_call": e0.elapsed_time(e1) / reps, "note": "B=128 sequences x gamma=6 chained probes, n=4"}
            # n-gram-assisted verify (config 4 shape: gamma=6, no drafter logits), greedy
            tn = sets[0][0][:128].repeat(1, 2, 1)[:, :7].contiguous()
            tkn = tn[:, :6].float().argmax(-1)
            tkn[::2, 3] = 7
            for i in range(2):
                sd.fused_verify(tn, None, tkn, None, None, seed=1, offset=i, greedy=True, flags=L.NGRAM)
            torch.cuda.synchronize()
            e0.record()
            for i in range(5):
                sd.fused_verify(tn, None, tkn, None, None, seed=1, offset=i, greedy=True, flags=L.NGRAM)
            e1.record()
            torch.cuda.synchronize()
            msn = e0.elapsed_time(e1) / 5
            sweep["ngram_verify_B128_g6_greedy"] = {"tokens_per_s": 128 * 6 / (msn * 1e-3), "ms_per_step": msn}
            del tn
        except Exception as ex:  # secondary numbers must never cost the headline line
            sweep["secondary_error"] = str(ex)[:200]

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp) and dtype == "bf16" and (B, g, V) == (256, 4, 128256):
            try:
                traffic = float(json.load(open(tp))["rowfast_tma_kernel_dram_bytes_per_launch"])
            except Exception:
                traffic = None
        ab = alg_bytes(B, g, V, dtype)
        ach = ab / (t_rowstats * 1e-3) / 1e9
        out = {
            "metric": "verified draft tokens/s (B=256, gamma=4, V=128k)", "value": value, "unit": "tokens/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config_dict(args, world),
            "roofline": {"bound": "hbm", "kernel": "rowfast_tma_kernel" if dtype != "f32" else "rowfast_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic,
                         "traffic_source": "ncu --set full capture of one full-batch launch (profiles/r2_traffic.json), dram read+write bytes",
                         "peak_source": peak_src,
                         "alg_bytes_per_launch": ab, "kernel_ms": t_rowstats, "decide_kernel_ms": t_decide,
                         "kernel_timed": ("alone: separate pass with the two-chunk stream pipelining off, one launch "
                                          "reads all algorithmic bytes" if pipelined else "inside the timed steps"),
                         "row_kernels_span_in_step_ms": t_rowstats_in_step, "tail_after_last_row_kernel_ms": t_decide_in_step,
                         "split_timed": "instrumented passes outside the K timed steps (event hooks cost ~10 us per step)",
                         "step_frac": ab / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / Ke, "pinned_buffers_numa_node": numa},
            # kernels of ours inside the timed region
            "gpu_launches": int(round(K * per_step)), "gpu_launches_per_step": by_name, "gpu_launches_source": launches_src,
            "gather": gather_kind,
            "graph_replay_ms_per_step": ms_graph,
            "two_batches_in_flight_ms_per_step": ms_two,
            "clocks": clocks,
        }
        if strong is not None:
            out["strong_scaling"] = strong
        if sweep:
            out["sweep"] = sweep
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def cpu_reference_leg(args, steps, warmup, all_threads=True):
    """The reference's CPU path for the same workload, on a bounded sample (B_s sequences of the same shape), timed on
    this host's cores.  kind = "reference": oracle/_ref holds the UNMODIFIED reference files (oracle/make_ref.py) and
    oracle/ref_arm.py drives one speculative step per sequence through the reference's own speculative_generate
    (sampling/speculative_decoding.py:107-172) and LogitsProcessor classes; kind = "port" (only when oracle/_ref did not
    travel): oracle/torch_port.py, the torch-eager restatement checked bit for bit against the reference."""
    import torch
    from oracle import ref_arm, torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores if all_threads else 1)
    Bs = args.cpu_sample_B
    g, V = args.gamma, args.V
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[args.dtype]
    gen = torch.Generator().manual_seed(1234)
    t = (3.0 * torch.randn(Bs, g + 1, V, generator=gen))
    d = (t[:, :g] + args.sigma * torch.randn(Bs, g, V, generator=gen)).to(dt)
    t = t.to(dt)
    mode = MODES[args.mode]
    kind = "reference" if ref_arm.available() else "port"
    if kind == "reference":
        # the reference's models emit fp32 logits on CPU: same VALUES as the GPU arm's rows, upcast outside the timing
        tf, df = t.float(), d.float()
        proc = ref_arm.make_processor(mode)
        one = lambda b: ref_arm.verify_step(tf[b], df[b], mode, proc)
    else:
        toks = torch.stack([torch_port.sample_rows(d[b], mode, gen) for b in range(Bs)])
        one = lambda b: torch_port.verify_one(t[b], d[b], toks[b], mode, gen)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for b in range(Bs):  # the reference is batch-1: one sequence per call
            one(b)
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    ms = 1e3 * sum(times) / len(times)
    val = Bs * g / (ms * 1e-3)
    what = ("unmodified reference (oracle/_ref: speculative_generate + LogitsProcessor, fp32 logits)" if kind == "reference"
            else "torch-eager port of the reference arithmetic (oracle/torch_port.py)")
    return {"value": val, "ms_sample": ms, "ms_full_batch": ms * args.B / Bs, "cores": cores, "kind": kind, "timed_steps": len(times),
            "sample": f"{Bs} of {args.B} sequences per step (same shape and values), {len(times)} timed steps, {what}, "
                      f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="multinomial", choices=list(MODES))
    ap.add_argument("--dtype", default="bf16", choices=list(ESIZE))
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--gamma", type=int, default=4)
    ap.add_argument("--V", type=int, default=128256)
    ap.add_argument("--sigma", type=float, default=0.5)
    ap.add_argument("--nbuf", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-B", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: all-gather the packed results with NCCL instead of P2P stores")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: B sequences per GPU (default); strong: B is the global batch, sharded across the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        # each "step" = the bounded sample (cpu_sample_B sequences); as many steps as asked for, capped so that the
        # whole run stays within a few minutes; `steps` reports what was actually timed, `ms_per_step` is scaled to
        # the full batch of B sequences (the sample is B_s / B of a step)
        r = cpu_reference_leg(args, max(1, min(args.steps, 20)), max(1, min(args.warmup, 2)))
        print(json.dumps({
            "impl": "reference", "metric": "verified draft tokens/s (B=256, gamma=4, V=128k)", "value": r["value"],
            "unit": "tokens/s", "n_gpus": args.gpus, "steps": r["timed_steps"], "steps_requested": args.steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["ms_full_batch"], "ms_per_sample_step": r["ms_sample"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    out = run_ours(args)
    if rank == 0 and out is not None:
        if args.gpus == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_leg(args, 3, 1)
                out["cpu_baseline"] = {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as e:  # the baseline is a reported number; never lose the GPU line over it
                out["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"failed: {e}"}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
