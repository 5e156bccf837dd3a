"""BASELINE.json configs[4]: synthetic stress sweep of the fine-stage train step (VoxurfF fwd+bwd) on one GPU:
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

dev = torch.device("cuda", 0)
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
        for n in ns:
            rays = {k: v.to(dev) for k, v in S.make_rays(n, 1234).items()}

            def step():
                for p in params:
                    p.grad = None
                m0 = m1 = m3 = 0
                for lo in range(0, n, TILE):
                    b = {k: v[lo:lo + TILE] for k, v in rays.items()}
                    out = model(s_val=20.0, **b)
                    (bench.loss_fn(out, b["rgbs"]) * (b["rays_o"].shape[0] / n)).backward()
                    st = model.last_streams["streams"]
                    m0 += int(st.cnt_inbox.sum()); m1 += st.m1; m3 += st.m3
                return m0, m1, m3

            for _ in range(3 if n <= (1 << 18) else 1):   # the second step after a model build can stall (bench.py)
                step()
            torch.cuda.synchronize()
            reps = 3 if n <= (1 << 18) else 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                m0, m1, m3 = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row = dict(grid=R, mask="sparse" if sparse else "dense", rays=n, ms_per_step=round(ms, 3),
                       rays_per_s=round(n / ms * 1e3), candidates_per_ray=round(m0 / n, 1), M1=m1, M3=m3,
                       candidate_samples_per_s=round(m0 / ms * 1e3), shaded_samples_per_s=round(m3 / ms * 1e3))
            rows.append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
        del model, params
        torch.cuda.empty_cache()
print(json.dumps({"workload": "VoxurfF fwd+bwd (bf16 tcgen05 MLPs), 1 B200, rays tiled by 2^16 per renderer call, s_val 20, "
                              "synthetic sphere scene; timing: CUDA events, 3 (1 for >= 2^20 rays) warm-up steps, 3 (1) timed steps",
                  "rows": rows}, indent=1))
