"""Builds the tracked round-2 summaries under profiles/ from the scratch files a `scripts/final_artifacts_r2.sh` run left in
gpurun_out/ (bench JSON lines, ncu launch list, ncu --set full raw pages) plus a SASS excerpt of the shipped library."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def load(f):
    try:
        return json.loads(open(os.path.join(G, f)).read().strip().splitlines()[-1])
    except Exception:
        return None


for f in os.listdir(G):
    if f.startswith("r2_bench_") and f.endswith(".json") or f.startswith("r2_scale_") and f.endswith(".json") \
            or f in ("r2_launches.csv", "r2_ncu_full_hot.csv", "r2_ncu_full_rowsel.csv", "r2_small_launches.csv"):
        shutil.copyfile(os.path.join(G, f), os.path.join(P, f))

# ---- launch list summary
rows = list(csv.reader(open(os.path.join(G, "r2_launches.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
d = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) > vi:
        d.setdefault((r[ki].split("(")[0].replace("void specdec::", ""), r[gi]), []).append(float(r[vi].replace(",", "")) / 1000)
with open(os.path.join(P, "r2_launches_summary.txt"), "w") as o:
    o.write("ncu --metrics gpu__time_duration.sum --clock-control none, `bench.py --steps 5 --warmup 3 --no-sweep` (bf16, B=256, gamma=4, V=128256)\n"
            "per-launch device time, cold-cache and serialised; grid 592 launches of 50 us = chunk 0 of a two-chunk step (128 sequences),\n"
            "grid 444 = chunk 1 (3 CTAs/SM), grid 592 at ~97 us = the FULL-BATCH launch of the single-chunk pass roofline.frac is computed from\n\n")
    for k, v in d.items():
        o.write(f"{k[0]:40s} grid {k[1]:16s} n={len(v):3d}  avg {sum(v)/len(v):7.1f} us  min {min(v):7.1f}  max {max(v):7.1f}\n")
    full = [x for (k, g), v in d.items() if "rowfast_tma" in k and g.startswith("(592") for x in v if x > 80]
    if full:
        ab = 591003648
        o.write(f"\nfull-batch rowfast_tma_kernel launches: {len(full)}, avg {sum(full)/len(full):.1f} us under ncu => {ab/ (sum(full)/len(full)) / 1e3:.0f} GB/s "
                f"(live CUDA-event timing of the same launch is in r2_bench_bf16.json:roofline.kernel_ms)\n")

# ---- ncu --set full summaries
with open(os.path.join(P, "r2_ncu_full_hot_kernels.txt"), "w") as o:
    o.write("ncu --set full --clock-control none, one launch each (bf16 B=256 gamma=4 V=128256, single-chunk step: the row kernel launch reads all 591 MB;\n"
            "rowsel_tma_kernel: the 1024-row launch of the drafter-side top-k 50 sample_rows call)\n\n")
    for f in ("r2_ncu_full_hot.csv", "r2_ncu_full_rowsel.csv"):
        rr = list(csv.reader(open(os.path.join(G, f))))
        hh = rr[0]
        for r in rr[2:]:
            o.write("Kernel " + r[hh.index("Kernel Name")] + "\n")
            for i, n in enumerate(hh):
                if "__" in n:
                    o.write(f"    {n} = {r[i]} {rr[1][i]}\n")
            o.write("\n")
    rr = list(csv.reader(open(os.path.join(G, "r2_ncu_full_hot.csv"))))
    hh = rr[0]
    for r in rr[2:]:
        if "rowfast_tma" in r[hh.index("Kernel Name")]:
            rd, wr = float(r[hh.index("dram__bytes_read.sum")]), float(r[hh.index("dram__bytes_write.sum")])
            json.dump({"rowfast_tma_kernel_dram_bytes_per_launch": (rd + wr) * 1e6, "read_mb": rd, "write_mb": wr,
                       "algorithmic_bytes": 591003648, "source": "profiles/r2_ncu_full_hot.csv (ncu --set full, full-batch launch, grid 592)"},
                      open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)

# ---- SASS excerpt: the instructions that prove the mechanisms (bulk copy + mbarrier in the row kernel, packed fp32 in the tail)
so = os.path.join(ROOT, "speculative-decoding_b200", "libspecdec_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, by = None, collections.defaultdict(list)
for line in sass.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        by[cur].append(line)
with open(os.path.join(P, "r2_sass_excerpt.txt"), "w") as o:
    o.write("cuobjdump -sass speculative-decoding_b200/libspecdec_b200.so (sm_100a), mnemonic counts per kernel and the first occurrences\n\n")
    for fn, pats in (("rowfast_tma_kernelILi1ELi0", ["UBLKCP", "SYNCS", "HMNMX2", "MUFU.EX2", "FFMA2", "FADD2", "LDS.128"]),
                     ("tail_slots_kernelILi1ELb0", ["FFMA2", "FADD2", "FMUL2", "F2I.U64", "LDS.128", "STS.128", "REDUX", "LD.E.64.STRONG", "NANOSLEEP"]),
                     ("rowsel_tma_kernelILi1ELb1ELb0", ["UBLKCP", "SYNCS", "HMNMX2", "BAR.SYNC"]),
                     ("plan_kernelILi1", ["MUFU.EX2", "ACQBULK", "LDG"])):
        names = [k for k in by if fn in k]
        for k in names[:1]:
            o.write(f"== {k}  ({len(by[k])} instructions)\n")
            for p in pats:
                hits = [l for l in by[k] if p in l]
                o.write(f"   {p:16s} x{len(hits)}\n")
                for l in hits[:2]:
                    o.write("        " + l.strip()[:150] + "\n")
            o.write("\n")
print("profiles written")

# ---- r2_results.md: every measured number of the round in one table
b = load("r2_bench_bf16.json")
lines = ["# Round-2 results (one B200 unless a row says otherwise; bf16, B=256, gamma=4, V=128256, multinomial T=1, sigma=0.5)", ""]
if b:
    r = b["roofline"]
    lines += ["| quantity | value |", "|---|---|",
              f"| verified draft tokens/s (device-resident inputs) | {b['value']/1e6:.2f} M ({b['ms_per_step']*1e3:.1f} us/step, {b['steps']} steps) |",
              f"| whole step as a fraction of the measured HBM peak ({r['peak']:.0f} GB/s) | {r['step_frac']:.3f} |",
              f"| dominant kernel `{r['kernel']}` alone | {r['kernel_ms']*1e3:.1f} us = {r['achieved']:.0f} GB/s = {r['frac']:.3f} of peak; DRAM traffic {r['traffic']/1e6 if r['traffic'] else float('nan'):.1f} MB for {r['alg_bytes_per_launch']/1e6:.1f} MB algorithmic |",
              f"| row kernels' span inside a step / tail after the last row kernel | {r['row_kernels_span_in_step_ms']*1e3:.1f} us / {r['tail_after_last_row_kernel_ms']*1e3:.1f} us |",
              f"| same step replayed from a CUDA graph | {b['graph_replay_ms_per_step']*1e3:.1f} us |",
              f"| two independent batches in flight on two streams | {(b.get('two_batches_in_flight_ms_per_step') or float('nan'))*1e3:.1f} us per step |",
              f"| e2e (pinned host logits -> H2D -> verify -> D2H) | {b['e2e']['value']/1e3:.1f} k tok/s, {b['e2e']['ms_per_step']:.2f} ms/step, {b['e2e']['h2d_bytes_per_step']/1e6:.0f} MB H2D |",
              f"| kernels of ours per step (CUPTI) | {b['gpu_launches_per_step']} |",
              f"| CPU baseline ({b['cpu_baseline']['kind']}, {b['cpu_baseline']['cores']} threads) | {b['cpu_baseline']['value']:.0f} tok/s |", ""]
    lines += ["| sweep entry | ms/step | tokens/s | step fraction of HBM peak | graph replay ms |", "|---|---|---|---|---|"]
    for k, v in b.get("sweep", {}).items():
        if "ms_per_step" in v:
            lines.append(f"| {k} | {v['ms_per_step']:.4f} | {v.get('tokens_per_s', float('nan'))/1e6:.3f} M | {v.get('step_frac_of_hbm_peak', float('nan')):.3f} | {v.get('graph_replay_ms_per_step', float('nan')):.4f} |")
        elif "ms_per_call" in v:
            lines.append(f"| {k} | {v['ms_per_call']:.4f} (per call) | | | |")
    lines.append("")
for name in ("f32", "f16", "topk50", "nucleus"):
    x = load(f"r2_bench_{name}.json")
    if x:
        lines.append(f"* `bench.py` {name}: {x['value']/1e6:.2f} M tok/s, {x['ms_per_step']*1e3:.1f} us/step, step_frac {x['roofline']['step_frac']:.3f}, kernel frac {x['roofline']['frac']:.3f}")
x = load("r2_bench_reference.json")
if x:
    lines.append(f"* `bench.py --impl reference`: {x['value']:.0f} tok/s ({x['cpu_baseline']['kind']}, {x['cpu_baseline']['cores']} threads, {x['steps']} timed steps, full-batch step {x['ms_per_step']:.0f} ms)")
sc = [(n, load(f"r2_scale_{n}gpu.json")) for n in (1, 2, 4, 8)]
if any(v for _, v in sc):
    lines += ["", "| GPUs | weak: tokens/s (B=256 per GPU) | ms/step | e2e tokens/s | strong: tokens/s (global B=256) | strong ms/step |", "|---|---|---|---|---|---|"]
    for n, v in sc:
        if v:
            st = v.get("strong_scaling") or {}
            lines.append(f"| {n} | {v['value']/1e6:.2f} M | {v['ms_per_step']*1e3:.1f} us | {v['e2e']['value']/1e3:.1f} k | "
                         f"{(st.get('value') or float('nan'))/1e6:.2f} M | {(st.get('ms_per_step') or float('nan'))*1e3:.1f} us |")
open(os.path.join(P, "r2_results.md"), "w").write("\n".join(lines) + "\n")
print("r2_results.md written")
