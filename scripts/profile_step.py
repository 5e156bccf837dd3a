"""torch.profiler view of one step of a bench workload (GPU box): which torch-side kernels and gaps surround the
library's kernels.   python scripts/profile_step.py [fine|lts|eval] [extra bench.py flags]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench

stage = sys.argv[1] if len(sys.argv) > 1 else "fine"
sys.argv = [sys.argv[0], "--stage", stage] + sys.argv[2:]
a = bench.parse()
dev = torch.device("cuda", 0)
model, host, fwd_kw, stage_loss = bench.build_stage(a, dev, 0)
params = [p for p in model.parameters() if p.requires_grad]
batch = {k: v.to(dev) for k, v in host.items()}

def step():
    if stage == "eval":
        n = batch["rays_o"].shape[0]
        for lo in range(0, n, a.eval_chunk):
            sl = slice(lo, min(lo + a.eval_chunk, n))
            model(rays_o=batch["rays_o"][sl], rays_d=batch["rays_d"][sl], viewdirs=batch["viewdirs"][sl],
                  em_modes=torch.tensor(0), **fwd_kw)
        return
    for p in params:
        p.grad = None
    out = model(**fwd_kw, **batch)
    stage_loss(out, batch["rgbs"]).backward()

for _ in range(6):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"steady state: {e0.elapsed_time(e1) / 3:.3f} ms/step (no profiler)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
print("==== host side (3 steps) ====")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU busy {busy/3e3:.3f} ms/step, span {span/3e3:.3f} ms/step, idle {(span-busy)/3e3:.3f} ms/step")
gaps = sorted(((ev[i+1].time_range.start - ev[i].time_range.end, ev[i].name[:50], ev[i+1].name[:50]) for i in range(len(ev)-1)), reverse=True)[:12]
for g in gaps: print(f"gap {g[0]:8.1f} us after {g[1]} before {g[2]}")

# full GPU timeline of the middle step (start us, duration us, idle gap before, name): where the device waits for the host
import json as _json
marks = [i for i, e in enumerate(ev) if "k_march<false>" in e.name or "k_march<0>" in e.name]
if len(marks) >= 3:
    lo, hi = marks[1], marks[2]
    t0 = ev[lo].time_range.start
    rows = []
    for i in range(lo, hi):
        e = ev[i]
        rows.append((e.time_range.start - t0, e.time_range.end - e.time_range.start,
                     e.time_range.start - ev[i - 1].time_range.end, e.name[:70]))
    out = os.path.join(ROOT, "gpurun_out", f"timeline_{stage}.txt")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        for r in rows:
            f.write(f"{r[0]:9.1f} {r[1]:8.1f} {r[2]:7.1f}  {r[3]}\n")
    print(f"timeline: {len(rows)} GPU events, idle {sum(max(r[2], 0) for r in rows) / 1e3:.3f} ms -> {out}")
