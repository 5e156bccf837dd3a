"""torch.profiler timeline of one multi-GPU step (backward tail + gradient exchange) on rank 0:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_exchange.py [bench flags]
writes gpurun_out/timeline_exchange_n<N>.txt: start us, duration us, stream, kernel name (from the tone-map backward on)."""
import datetime
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist

import bench
from esr_nerf_b200.dist import GridGradCompactor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
sys.argv = [sys.argv[0], "--gpus", str(world)] + sys.argv[1:]
a = bench.parse()
model, host, fwd_kw, stage_loss = bench.build_stage(a, dev, rank)
params = [p for p in model.parameters() if p.requires_grad]
batch = {k: v.to(dev) for k, v in host.items()}
comp = GridGradCompactor(model)
if os.environ.get("ESR_ALLREDUCE_OVERLAP", "1") not in ("0", ""):
    comp.overlap_color_allreduce(True)


def step():
    for p in params:
        p.grad = None
    out = model(**fwd_kw, **batch)
    stage_loss(out, batch["rgbs"]).backward()
    comp.allreduce()


for _ in range(8):
    step()
torch.cuda.synchronize()
dist.barrier()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    marks = [i for i, e in enumerate(ev) if "k_march_count" in e.name or "k_march<" in e.name]
    firsts = [m for j, m in enumerate(marks) if j == 0 or ev[m].time_range.start - ev[marks[j - 1]].time_range.start > 5000]
    lo, hi = firsts[1], firsts[2]
    t0 = ev[lo].time_range.start
    out = os.path.join(ROOT, "gpurun_out", f"timeline_exchange_n{world}.txt")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    started = False
    with open(out, "w") as f:
        f.write(f"# step span {(ev[hi].time_range.start - t0) / 1e3:.3f} ms\n")
        for e in ev[lo:hi]:
            started = started or "tonemap_bwd" in e.name
            if started:
                f.write(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}\n")
    print(open(out).read()[-6000:])
dist.barrier()
dist.destroy_process_group()
