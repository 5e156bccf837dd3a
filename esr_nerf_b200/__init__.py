"""esr_nerf_b200 — B200-native (sm_100a) render hot path of ESR-NeRF behind the reference's own
render-function signatures.  See DESIGN.md / INTEGRATION.md.

Public surface (mirrors the reference names):
    ESRNeRF                               <- app.fine.model.ESRNeRF   (LTS / PDRA stages; forward_training)
    VoxurfF                               <- app.fine.model.VoxurfF
    VoxurfC                               <- app.coarse.model.VoxurfC
    DVGO                                  <- app.coarse.model.DVGO
    render_utils.render_utils_cuda.*      <- app/utils/base/cuda/render_utils.cpp (live entries)
    render_utils.total_variation_cuda.*   <- app/utils/base/cuda/total_variation.cpp
    render_utils.segment_coo              <- torch_scatter.segment_coo(reduce="sum")
    render_utils.Alphas2Weights           <- app/utils/base/module.py:117-143
    samplers.BatchSampler / RayGroupManager <- utils2/utils.py:41-312 (index-only shuffles, rank slices)
    optimizer.*                           <- app/utils/optimizer.py (fused Adam step)
"""
from . import _lib  # noqa: F401
from ._lib import EsrError, build  # noqa: F401


def __getattr__(name):  # lazy: importing the package must work on a box without a GPU
    if name == "VoxurfF":
        from .voxurff import VoxurfF
        return VoxurfF
    if name == "ESRNeRF":
        from .esrnerf import ESRNeRF
        return ESRNeRF
    if name == "DVGO":
        from .dvgo import DVGO
        return DVGO
    if name == "VoxurfC":
        from .voxurfc import VoxurfC
        return VoxurfC
    if name in ("render_utils", "fused", "modules", "synthetic", "voxurff", "voxurfc", "dvgo", "dist", "esrnerf", "pbr", "samplers", "optimizer"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
