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
    cot = {k: v.to(dev) for k, v in D.shard_batch(C.cotangents(n), rank, world).items()}
    params = [p for p in model.parameters() if p.requires_grad]

    def local_backward():
        for p in params:
            p.grad = None
        o = model(s_val=float(fx["s_val"]), **batch)
        sum((o[k] * cot[k]).sum() for k in cot).backward()

    # reference: dense all-reduce of every gradient
    fused.COLOR_GRADS_READY_HOOK = None
    local_backward()
    D.allreduce_gradients(params)
    ref = [p.grad.clone() for p in params]
    worst = 0.0
    for overlap in (False, True):
        comp = D.GridGradCompactor(model)
        comp.overlap_color_allreduce(overlap)
        local_backward()
        assert (comp._early is not None) == overlap
        comp.allreduce(verify=not overlap)
        assert comp._early is None
        for p, r in zip(params, ref):
            assert torch.equal(p.grad, r), (overlap, tuple(p.shape))
            worst = max(worst, float((p.grad - r).abs().max()))
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
    assert out.get(timeout=5) == 0.0
