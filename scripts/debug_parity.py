"""Debug helper (GPU box): print every stream / output / gradient error of the fused path vs golden + oracle port."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import esr_testlib as C
from esr_nerf_b200 import synthetic as S
from oracle import voxurf_port as P

DEV = "cuda:0"
def run(case, mode, on_first, n=None, seed=None):
    fx, weights = C.load_case(case)
    n = n or int(fx["n_rays"]); seed = seed or int(fx["ray_seed"])
    rays = S.make_rays(n, seed)
    scene = C.oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    params, leaves = C.oracle_params(scene, weights)
    ref, inter = P.voxurff_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"], rays["em_modes"], float(fx["s_val"]))
    cot = C.cotangents(n)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    m = C.build_product_model(fx, weights, DEV)
    m.mlp_mode, m.on_first_order, m.keep_streams = mode, on_first, True
    out = m(s_val=float(fx["s_val"]), **{k: v.to(DEV) for k, v in rays.items()})
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    st = m.last_streams["streams"]
    print(f"== {case} mode={mode} on_first={on_first} n={n}: m0 {int(st.cnt_inbox.sum())} vs {inter['m0']}; m1 {st.m1} vs {inter['m1_ray'].numel()}; m3 {st.m3} vs {inter['m3_ray'].numel()}")
    s_ray, s_step = st.s_ray.long().cpu(), st.s_step.long().cpu()
    o1 = torch.argsort(s_ray * (1 << 20) + s_step, stable=True)
    if st.m1 == inter["m1_ray"].numel():
        print("   m1 ray eq", torch.equal(s_ray[o1], inter["m1_ray"]), "step eq", torch.equal(s_step[o1], inter["m1_step"]),
              "sdf err", C.rel_err(st.s_sdf.cpu()[o1], inter["m1_sdf"]), "alpha err", C.rel_err(st.s_alpha.cpu()[o1], inter["m1_alpha"]),
              "sdf bit-equal", torch.equal(st.s_sdf.cpu()[o1], inter["m1_sdf"].detach()))
    ray, step, w = st.h_ray.long().cpu(), st.h_step.long().cpu(), m.last_streams["h_w"].cpu()
    o3 = torch.argsort(ray * (1 << 20) + step, stable=True)
    if st.m3 == inter["m3_ray"].numel():
        print("   m3 ray eq", torch.equal(ray[o3], inter["m3_ray"]), "step eq", torch.equal(step[o3], inter["m3_step"]), "w err", C.rel_err(w[o3], inter["m3_weights"]),
              "w bit-equal", torch.equal(w[o3], inter["m3_weights"].detach()))
        print("   lin err", C.rel_err(m.last_streams["lin"].cpu()[o3], inter["m3_lin"]), "rgb err", C.rel_err(m.last_streams["rgb"].cpu()[o3], inter["m3_rgb"]))
    for k in ref:
        print(f"   out {k:20s} {C.rel_err(out[k], ref[k]):.3e}")
    for name, p in m.named_parameters():
        if name in leaves and leaves[name].grad is not None:
            g = p.grad.contiguous() if p.grad is not None else None
            if g is None:
                print(f"   grad {name:32s} MISSING"); continue
            r = leaves[name].grad
            d = (g.cpu().double() - r.double()).abs()
            print(f"   grad {name:32s} rel {C.rel_err(g, r):.3e}  |ref|max {r.abs().max():.3e} sum-rel {abs(g.double().sum().item()-r.double().sum().item())/max(r.double().abs().sum().item(),1e-12):.3e} nnz {int((g!=0).sum())} vs {int((r!=0).sum())}")

if __name__ == "__main__":
    torch.manual_seed(0)
    run("fine_sparse_s20", "torch_fp32", False)
    run("fine_sparse_s20", "torch_fp32", True)
    run("fine_sparse_s20", "bf16", True)
    run("fine_dense_s220", "bf16", True)
    run("fine_sparse_s60_big", "torch_fp32", True, 2048, 31337)
    run("fine_sparse_s60_big", "bf16", True, 2048, 31337)
