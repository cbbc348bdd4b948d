"""Timing probe: eager vs CUDA-graph replay of the headline verify step (is the step host-bound?)."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import specdec_b200 as sd
B, g, V = 256, 4, 128256
gen = torch.Generator(device="cuda").manual_seed(0)
sets = []
for i in range(3):
    t = (3.0 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
    d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
    toks = sd.sample_rows(d.reshape(B * g, V), None, seed=4321, offset=0, seq_id0=0)[0].reshape(B, g)
    sets.append((t, d, toks))
def step(i):
    t, d, k = sets[i % 3]
    return sd.fused_verify(t, d, k, None, None, seed=1, offset=7)
for i in range(5): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 300
t0 = time.perf_counter(); e0.record()
for i in range(N): step(i)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"eager: {e0.elapsed_time(e1)/N:.4f} ms/step on device, host enqueue {1e3*(t1-t0)/N:.4f} ms/step")
graphs = []
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3): step(i)
torch.cuda.current_stream().wait_stream(s)
for i in range(3):
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        r = step(i)
    graphs.append((gph, r))
for i in range(6): graphs[i % 3][0].replay()
torch.cuda.synchronize()
e0.record()
for i in range(N): graphs[i % 3][0].replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/N:.4f} ms/step")
