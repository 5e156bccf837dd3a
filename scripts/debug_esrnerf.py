"""Diagnostics: product ESRNeRF vs the oracle port on one golden case (prints every comparison, asserts nothing)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import esr_testlib as C
from esr_nerf_b200 import synthetic as S
from oracle import esrnerf_port as E

DEV = "cuda:0"
case = sys.argv[1] if len(sys.argv) > 1 else "lts_sparse_s220"
fx, weights = C.load_esrnerf_case(case)
m = C.build_product_esrnerf(fx, weights, DEV)
m.keep_streams = True
m.draws = E.FixedDraws(int(fx["draw_seed"]))
n = int(fx["n_rays"])
rays = S.make_rays(n, int(fx["ray_seed"]))
batch = {k: v.to(DEV) for k, v in rays.items() if k != "rgbs"}
out = m(s_val=float(fx["s_val"]), uncert_masks=S.uncert_masks(n).to(DEV), normal_eps=float(fx["normal_eps"]),
        emit_eps=float(fx["emit_eps"]), **batch)
ref, inter, leaves, _ = C.run_esrnerf_port(fx, weights)
st = m.last_streams["streams"]
pts = m.last_streams["pts"].cpu()
d = (pts != inter["m3_pts"])
print("pts mismatching entries", int(d.sum()), "of", d.numel(), "max abs diff", float((pts - inter["m3_pts"]).abs().max()))
if d.any():
    i = torch.nonzero(d)[0]
    print("first", i.tolist(), pts[i[0]].tolist(), inter["m3_pts"][i[0]].tolist(), "ray/step", int(st.h_ray[i[0]]), int(st.h_step[i[0]]))
print("s_sdf equal", torch.equal(st.s_sdf.cpu(), inter["m1_sdf"].detach()), C.rel_err(st.s_sdf, inter["m1_sdf"]))
st2 = m.last_streams["lts"]["streams"]
print("lts m3", st2.m3, inter["lts"]["m3_ray"].shape[0],
      "ray eq", st2.m3 == inter["lts"]["m3_ray"].shape[0] and torch.equal(st2.h_ray.long().cpu(), inter["lts"]["m3_ray"]))
for k in sorted(out):
    if tuple(out[k].shape) != tuple(ref[k].shape):
        print(k, "SHAPE", tuple(out[k].shape), tuple(ref[k].shape))
    else:
        print(f"{k:20s} rel_err {C.rel_err(out[k], ref[k]):.3e}")
cot = C.esrnerf_cotangents(out)
loss = sum((out[k] * cot[k].to(DEV)).sum() for k in cot)
loss.backward()
l2 = sum((ref[k] * cot[k]).sum() for k in cot)
l2.backward()
print("loss", loss.item(), l2.item(), float(fx["loss"]))
for name, p in m.named_parameters():
    if name in leaves and leaves[name].grad is not None:
        if p.grad is None:
            print(name, "NO GRAD"); continue
        mx, l2e = C.grad_err(p.grad.contiguous(), leaves[name].grad)
        print(f"{name:32s} max {mx:.3e} l2 {l2e:.3e}")
