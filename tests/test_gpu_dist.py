"""2-GPU NCCL test (-m gpu, skipped on a single-GPU box) of the gradient exchange of the ray-sharded fine-stage step
(SURVEY.md §8e): the occupancy-compacted all-reduce — with and without the colour volumes' part started from inside the
backward pass (fused.COLOR_GRADS_READY_HOOK) — must equal the dense all-reduce of every gradient, bit for bit on two
ranks (a two-term float sum is commutative)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist

    import esr_testlib as C
    from esr_nerf_b200 import dist as D
    from esr_nerf_b200 import fused
    from esr_nerf_b200 import synthetic as S

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    fx, weights = C.load_case("fine_sparse_s60_big")
    model = C.build_product_model(fx, weights, str(dev))
    n = 4096
    batch = {k: v.to(dev) for k, v in D.shard_batch(S.make_rays(n, 21), rank, world).items()}
    sl = D.shard_slice(n, rank, world)
    cot = {k: v[sl].to(dev) for k, v in C.cotangents(n).items()}
    params = [p for p in model.parameters() if p.requires_grad]

    def local_backward():
        for p in params:
            p.grad = None
        o = model(s_val=float(fx["s_val"]), **batch)
        sum((o[k] * cot[k]).sum() for k in cot).backward()

    # local gradients once (the kernels accumulate with float REDs: a second backward differs in the last bits)
    fused.COLOR_GRADS_READY_HOOK = None
    local_backward()
    local = [p.grad.clone() for p in params]
    # reference: dense all-reduce of every gradient
    D.allreduce_gradients(params)
    ref = [p.grad.clone() for p in params]
    comp = D.GridGradCompactor(model)
    color = comp.grids[1:]
    worst = 0.0
    for overlap in (False, True):   # same local gradients through the compacted exchange: bit-exact
        for p, g in zip(params, local):
            p.grad = None if (overlap and any(p is c for c in color)) else g.clone()
        if overlap:                 # what Shade.backward does when the colour gradients are final
            bufs = {p: g.clone() for p, g in zip(params, local) if any(p is c for c in color)}
            comp._on_color_grads(bufs)
            assert comp._early is not None
            for p in color:
                p.grad = bufs[p]
        comp.allreduce(verify=True)
        assert comp._early is None
        for p, r in zip(params, ref):
            assert torch.equal(p.grad, r), (overlap, tuple(p.shape))
    # and inside autograd: the hook fires from the backward pass, the result matches within the REDs' reordering
    comp.overlap_color_allreduce(True)
    local_backward()
    assert comp._early is not None
    comp.allreduce()
    for p, r in zip(params, ref):
        err = float((p.grad - r).abs().max()) / (float(r.abs().max()) + 1e-30)
        assert err < 1e-3, (tuple(p.shape), err)      # bf16-MLP gradients: RED order + tensor-core split-K order
        worst = max(worst, err)
    fused.COLOR_GRADS_READY_HOOK = None
    if rank == 0:
        out.put(worst)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_compacted_and_overlapped_allreduce_equal_dense_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-3
