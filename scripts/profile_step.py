"""torch.profiler view of one training step of the bench workload (GPU box): which torch-side kernels and gaps
surround the library's kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from esr_nerf_b200 import synthetic as S
from esr_nerf_b200.voxurff import VoxurfF

dev = "cuda:0"
model = VoxurfF(S.fine_cfg(device=dev), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX,
                S.MASK_ALPHA_INIT, S.mask_density(100, True), 20.0, 256 ** 3)
model.load_state_dict({**model.state_dict(), **bench.random_mlp_weights()})
S.fill_fine_model(model)
params = [p for p in model.parameters() if p.requires_grad]
batch = {k: v.to(dev) for k, v in S.make_rays(1 << 16, 1234).items()}

def step():
    for p in params:
        p.grad = None
    out = model(s_val=20.0, **batch)
    bench.loss_fn(out, batch["rgbs"]).backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU busy {busy/3e3:.3f} ms/step, span {span/3e3:.3f} ms/step, idle {(span-busy)/3e3:.3f} ms/step")
gaps = sorted(((ev[i+1].time_range.start - ev[i].time_range.end, ev[i].name[:50], ev[i+1].name[:50]) for i in range(len(ev)-1)), reverse=True)[:15]
for g in gaps: print(f"gap {g[0]:8.1f} us after {g[1]} before {g[2]}")
