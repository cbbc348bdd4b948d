"""profiles/r1_results.md from the raw bench JSON lines in profiles/ (run after scripts/final_artifacts.sh and
scripts/scaling.sh have been copied there)."""
import json
import os

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles") + "/"


def load(f):
    return json.loads(open(P + f).read().strip().split("\n")[-1])


b, f32, f16 = load("r1_bench_bf16.json"), load("r1_bench_f32.json"), load("r1_bench_f16.json")
ref, refn = load("r1_bench_reference.json"), load("r1_bench_reference_nucleus0.9.json")
sc = {n: load(f"r1_scale_{n}gpu.json") for n in (1, 2, 4, 8)}


def row(name, d):
    r = d["roofline"]
    return (f"| {name} | {d['value']:.3g} | {d['ms_per_step']:.4f} | {d['graph_replay_ms_per_step']:.4f} | {r['kernel_ms']:.4f} | "
            f"{r['frac']:.3f} | {r['step_frac']:.3f} | {d['e2e']['value']:.3g} |")


L = ["# Round-1 measured results (B200, `gpurun`, one fresh box per call)\n",
     "Workload: synthetic logits verify, B=256, γ=4, V=128256, `3·randn` target, drafter = target + 0.5·randn, draft tokens "
     "drawn from the drafter (BASELINE.json configs[1]). Peak = MEASURED_PEAKS.json hbm_gbs = 6543.7 GB/s. Raw JSON lines: "
     "`r1_bench_*.json`, `r1_scale_*gpu.json`.\n",
     "| dtype / mode | verified draft tok/s | ms / step | ms / step, CUDA-graph replay | row kernel ms (timed alone) | row kernel "
     "frac of HBM peak | whole-step frac | e2e tok/s (host buffers) |",
     "|---|---|---|---|---|---|---|---|",
     row("bf16 multinomial T=1 (headline)", b), row("fp32 multinomial T=1", f32), row("fp16 multinomial T=1", f16),
     "\nRound history of the headline step: 0.2219 ms (round start: row kernel + plan + exact_rows + sample_partial, timed with "
     "the event hooks on) → 0.2164 (fused tail with the weights of the deciding row pair cached in shared memory + two-chunk "
     "stream pipelining) → 0.2031 (timed steps without the six event records of the per-kernel split) → 0.1968 (programmatic "
     "dependent launches) → 0.190 (row kernel: stage-level packed maximum) → 0.178 ms (early `griddepcontrol.launch_dependents`: "
     "the next kernel's CTAs are scheduled while the previous one drains).  Row kernel alone: 0.1095 ms (82.5 % of the "
     "measured HBM peak) → 0.1043 (3 × 16 KB TMA stages) → 0.096 ms (94 %).\n",
     "Secondary sweep (bf16, same inputs unless noted; whole step):\n",
     "| case | ms / step | tok/s | whole-step frac of HBM peak |", "|---|---|---|---|"]
for k, v in b["sweep"].items():
    if "ms_per_step" in v:
        L.append(f"| {k} | {v['ms_per_step']:.4f} | {v['tokens_per_s']:.3g} | {v.get('step_frac_of_hbm_peak', 0):.3f} |")
    else:
        L.append(f"| {k} | {v['ms_per_call']:.4f} (per call) | - | {v.get('note', '')} |")
L += ["\nTop-p 0.9 history: flat `3·randn` rows 2.297 ms (round start, exact band search) → 1.40 ms (`nucleus_hist_kernel`); "
      "near-uniform `0.05·randn` rows (what random-init models emit): 6.26 ms at B=64 (≈25 ms at B=256) → 0.48 ms at B=64, "
      "1.41 ms at B=256.\n",
      "CPU baseline / reference arm (`bench.py --impl reference`: `oracle/torch_port.py`, torch-eager restatement of the "
      "reference arithmetic, 16 host cores, 4 of 256 sequences per step):\n",
      "| mode | tok/s |", "|---|---|", f"| multinomial T=1 | {ref['value']:.1f} |", f"| nucleus 0.9 | {refn['value']:.2f} |",
      f"\nClocks during the timed region (NVML, 2 ms period): {json.dumps(b['clocks'])}\n",
      "Weak scaling on one 8×B200 box (torchrun, one rank per GPU, sequences sharded by rank, async all-gather of the packed "
      "int32 results over NCCL; measured on the 0.190 ms build, one commit before the early dependent launches):\n",
      "| GPUs | verified draft tok/s | ms / step (max over ranks) | × 1 GPU |", "|---|---|---|---|"]
for n in (1, 2, 4, 8):
    L.append(f"| {n} | {sc[n]['value']:.4g} | {sc[n]['ms_per_step']:.4f} | {sc[n]['value'] / sc[1]['value']:.2f} |")
open(P + "r1_results.md", "w").write("\n".join(L) + "\n")
print("\n".join(L))
