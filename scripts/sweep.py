"""BASELINE.json configs[4]: synthetic stress sweep of the fine-stage train step (VoxurfF fwd+bwd) on 1..8 GPUs
(`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep.py`: the rays of a step are
sharded over the ranks — strong scaling of the global ray count — and the gradients exchanged once per step):
rays per step 2^14..2^22 x grid 32^3..512^3 (= ~64..1024 candidate samples per ray at stepsize 0.5) x {dense, sparse}
MaskCache.  The caller tiles the rays (<= 2^16 per renderer call, gradients accumulate across tiles) so the candidate
stream of a step (up to 4.3e9 samples) is never materialised.

    python scripts/sweep.py [--quick] > profiles/<round>_sweep.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from esr_nerf_b200 import synthetic as S
from esr_nerf_b200.voxurff import VoxurfF

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import datetime
    import torch.distributed as dist
    from esr_nerf_b200.dist import GridGradCompactor, shard_slice
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
quick = "--quick" in sys.argv
TILE = 1 << 16
grids = [32, 64, 128, 256, 512]
ns = [1 << 14, 1 << 16, 1 << 18, 1 << 20] + ([] if quick else [1 << 22])
weights = bench.random_mlp_weights()
rows = []
for R in grids:
    for sparse in (False, True):
        model = VoxurfF(S.fine_cfg(device=str(dev)), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX,
                        S.MASK_ALPHA_INIT, S.mask_density(min(100, R), sparse), 20.0, R ** 3)
        model.load_state_dict({**model.state_dict(), **weights})
        S.fill_fine_model(model)
        model.keep_streams = True
        params = [p for p in model.parameters() if p.requires_grad]
        exchange = GridGradCompactor(model) if world > 1 else None
        for n_global in ns:
            rays = S.make_rays(n_global, 1234)
            if world > 1:
                sl = shard_slice(n_global, rank, world)
                rays = {k: v[sl] for k, v in rays.items()}
            rays = {k: v.to(dev) for k, v in rays.items()}
            n = rays["rays_o"].shape[0]

            def step():
                for p in params:
                    p.grad = None
                m0 = m1 = m3 = 0
                for lo in range(0, n, TILE):
                    b = {k: v[lo:lo + TILE] for k, v in rays.items()}
                    out = model(s_val=20.0, **b)
                    (bench.loss_fn(out, b["rgbs"]) * (b["rays_o"].shape[0] / n_global)).backward()
                    st = model.last_streams["streams"]
                    m0 += int(st.cnt_inbox.sum()); m1 += st.m1; m3 += st.m3
                if exchange is not None:
                    exchange.allreduce()
                return m0, m1, m3

            for _ in range(3 if n_global <= (1 << 18) else 1):   # the second step after a model build can stall (bench.py)
                step()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            reps = 3 if n_global <= (1 << 18) else 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                m0, m1, m3 = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            cnt = torch.tensor([ms, m0, m1, m3], dtype=torch.float64, device=dev)
            if dist is not None:      # step time: max over the ranks; sample counts: summed
                t = cnt[:1].clone()
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
                cnt[0] = t[0]
            ms, m0, m1, m3 = (float(v) for v in cnt.tolist())
            row = dict(grid=R, mask="sparse" if sparse else "dense", rays=n_global, n_gpus=world, ms_per_step=round(ms, 3),
                       rays_per_s=round(n_global / ms * 1e3), candidates_per_ray=round(m0 / n_global, 1), M1=int(m1), M3=int(m3),
                       candidate_samples_per_s=round(m0 / ms * 1e3), shaded_samples_per_s=round(m3 / ms * 1e3))
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), file=sys.stderr, flush=True)
        del model, params, exchange
        torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"workload": f"VoxurfF fwd+bwd (tcgen05 MLPs, mlp_mode x2), {world} B200, the step's rays sharded over the ranks "
                                  "(one compacted gradient exchange per step), tiled by 2^16 per renderer call, s_val 20, synthetic "
                                  "sphere scene; timing: CUDA events, max over ranks, 3 (1 for >= 2^20 rays) warm-up steps, 3 (1) timed steps",
                      "rows": rows}, indent=1))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
