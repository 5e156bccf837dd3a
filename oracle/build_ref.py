"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Builds the reference's OWN native extensions for sm_100a from the sources where they lie under
/root/reference (nothing is copied into the repo):

    app/utils/base/cuda/render_utils.cpp + render_utils_kernel.cu      -> oracle/_ref/ref_render_utils_cuda.so
    app/utils/base/cuda/total_variation.cpp + total_variation_kernel.cu -> oracle/_ref/ref_total_variation_cuda.so

The only accommodation is a forced-include shim (oracle/ref_dispatch_shim.h) that restores the old
AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...) calling convention for torch 2.11.  The .so files are
git-ignored but travel to the GPU box with the snapshot, where tests/test_gpu_oracle_pin.py uses them to
pin the C restatement (oracle/render_utils_ref.c) and the product kernels against the reference's real
kernels on identical inputs.  /root/reference itself does not exist on the GPU box: load() only imports
the prebuilt modules there.

    python -m oracle.build_ref          # build (this container)
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("ESR_REFERENCE_ROOT", "/root/reference")
CUDA_DIR = os.path.join(REF_ROOT, "app", "utils", "base", "cuda")
MODULES = {
    "ref_render_utils_cuda": ["render_utils.cpp", "render_utils_kernel.cu"],
    "ref_total_variation_cuda": ["total_variation.cpp", "total_variation_kernel.cu"],
}


def _so_path(name: str) -> str:
    return os.path.join(OUT, name, f"{name}.so")


def build(quiet: bool = False, force: bool = False) -> None:
    if not os.path.isdir(CUDA_DIR):
        raise RuntimeError(f"{CUDA_DIR} not present (the reference does not travel to the GPU box)")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    shim = os.path.join(HERE, "ref_dispatch_shim.h")
    for name, files in MODULES.items():
        srcs = [os.path.join(CUDA_DIR, f) for f in files]
        so = _so_path(name)
        if not force and os.path.isfile(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs + [shim]):
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        # is_python_module=False: only build here; importing would need a CUDA runtime context on some ops
        load(name=name, sources=srcs, build_directory=bdir, verbose=not quiet, is_python_module=False,
             extra_cflags=["-include", shim, "-w"],
             extra_cuda_cflags=["-include", shim, "-w", "-gencode", "arch=compute_100a,code=sm_100a"])
        if not quiet:
            print(f"[build_ref] built {so}")


def available() -> bool:
    return all(os.path.isfile(_so_path(n)) for n in MODULES)


def load_module(name: str):
    """Import a prebuilt reference extension (pybind11 module) from oracle/_ref."""
    import torch  # noqa: F401  (libtorch must be loaded first)

    so = _so_path(name)
    if not os.path.isfile(so):
        raise RuntimeError(f"{so} missing: run `python -m oracle.build_ref` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)
