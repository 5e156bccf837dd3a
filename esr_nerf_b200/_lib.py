"""Build + ctypes binding of libesr_b200.so (the C ABI declared in include/esr_b200.h).

The library is built IN-TREE with nvcc for sm_100a only and loaded with ctypes; there is no
CPU fallback and no other backend: if the shared object is missing and cannot be built the
import of any render op raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import sys
from typing import Optional

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libesr_b200.so")
SOURCES = ["native_ops.cu", "voxurf_stream.cu", "encode.cu", "mlp_api.cu", "mlp_tc.cu", "dvgo.cu", "esrnerf.cu", "optim.cu", "grad_exchange.cu",
           "regularizers.cu", "render_step.cu"]
HEADERS = ["common.cuh", "scan.cuh", "mlp_layout.cuh", os.path.join("..", "..", "include", "esr_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

_lib: Optional[ctypes.CDLL] = None


class EsrError(RuntimeError):
    pass


def _nvcc() -> Optional[str]:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.isfile(cand) else None


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into esr_nerf_b200/libesr_b200.so.  Safe under torchrun: the stale check and
    the build run under an exclusive file lock, objects and the library are written to temporary names and renamed into
    place, so a rank never links or loads a half-written file (the ranks that waited find the library fresh)."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl

    os.makedirs(os.path.join(PKG_DIR, "build"), exist_ok=True)
    with open(os.path.join(PKG_DIR, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another process built it while this one waited for the lock
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.isfile(LIB_PATH):
            return LIB_PATH  # prebuilt library travelled with the snapshot; no compiler on this box
        raise EsrError("libesr_b200.so is missing and nvcc was not found; the B200 render path has no fallback")
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(build_dir, s.replace(".cu", f".{os.getpid()}.o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise EsrError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp_lib = f"{LIB_PATH}.{os.getpid()}.tmp"
    cmd = [nvcc, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for o in objs:
        if os.path.isfile(o):
            os.remove(o)
    if r.returncode != 0:
        raise EsrError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, LIB_PATH)
    print(f"[esr_nerf_b200] compiled {len(SOURCES)} CUDA sources for sm_100a -> {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


class Scene(ctypes.Structure):
    """esr_scene_t"""
    _fields_ = [
        ("xyz_min", ctypes.c_float * 3), ("xyz_max", ctypes.c_float * 3),
        ("gx", ctypes.c_int32), ("gy", ctypes.c_int32), ("gz", ctypes.c_int32),
        ("mask_xyz_min", ctypes.c_float * 3), ("mask_xyz_max", ctypes.c_float * 3),
        ("mx", ctypes.c_int32), ("my", ctypes.c_int32), ("mz", ctypes.c_int32),
        ("near", ctypes.c_float), ("far", ctypes.c_float),
        ("stepdist", ctypes.c_float), ("voxel_size", ctypes.c_float),
        ("act_shift", ctypes.c_float), ("mask_thres", ctypes.c_float),
        ("fast_thres", ctypes.c_float), ("s_val", ctypes.c_float), ("alpha_thres", ctypes.c_float),
        ("fd_eps", ctypes.c_float), ("sdf_tap_manual", ctypes.c_int32),
    ]


class DvgoScene(ctypes.Structure):
    """esr_dvgo_scene_t"""
    _fields_ = [("xyz_min", ctypes.c_float * 3), ("xyz_max", ctypes.c_float * 3),
                ("gx", ctypes.c_int32), ("gy", ctypes.c_int32), ("gz", ctypes.c_int32),
                ("near", ctypes.c_float), ("far", ctypes.c_float), ("stepdist", ctypes.c_float),
                ("interval", ctypes.c_float), ("act_shift", ctypes.c_float)]


class MlpDesc(ctypes.Structure):
    """esr_mlp_desc_t"""
    _fields_ = [("k0", ctypes.c_int32), ("width", ctypes.c_int32), ("n_hidden", ctypes.c_int32),
                ("n_out", ctypes.c_int32), ("act", ctypes.c_int32), ("precision", ctypes.c_int32)]


class VoxurffStep(ctypes.Structure):
    """esr_voxurff_step_t"""
    _fields_ = [("scene", Scene), ("mask_density", ctypes.c_void_p), ("mask_cls", ctypes.c_void_p),
                ("sdf_grid", ctypes.c_void_p), ("off_color_grid", ctypes.c_void_p), ("emo_color_grid", ctypes.c_void_p),
                ("flat_off", ctypes.c_void_p), ("flat_emo", ctypes.c_void_p), ("flat_tone", ctypes.c_void_p),
                ("precision", ctypes.c_int32), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
                ("n_rays", ctypes.c_int64), ("n_on", ctypes.c_int64), ("m1", ctypes.c_int64), ("m3", ctypes.c_int64),
                ("m3_on", ctypes.c_int64), ("workspace_used", ctypes.c_int64), ("workspace_needed", ctypes.c_int64),
                ("alphainv_last", ctypes.c_void_p), ("slot", ctypes.c_void_p * 32)]


P = ctypes.c_void_p
I64 = ctypes.c_int64
I32 = ctypes.c_int
F32 = ctypes.c_float
F3 = ctypes.c_float * 3
F3P = ctypes.POINTER(ctypes.c_float)
SCENE_P = ctypes.POINTER(Scene)
DESC_P = ctypes.POINTER(MlpDesc)
DVGO_P = ctypes.POINTER(DvgoScene)

# every symbol declared in include/esr_b200.h: name -> (restype, argtypes)
PROTOTYPES = {
    "esr_last_error": (ctypes.c_char_p, []),
    "esr_version": (I32, []),
    "esr_launch_count": (I64, []),
    "esr_stage_timing": (I32, [I32]),
    "esr_stage_timing_report": (I64, [ctypes.c_char_p, I64]),
    "esr_scan_scratch_bytes": (I64, [I64]),
    "esr_sample_pts_on_rays_count": (I32, [P, P, F3P, F3P, F32, F32, F32, I64, P, P, P, P, P, P, P]),
    "esr_sample_pts_on_rays_fill": (I32, [P, P, F3P, F3P, F32, F32, F32, I64, P, I64, P, P, P, P, P]),
    "esr_alpha2weight_fwd": (I32, [P, P, I64, I64, P, P, P, P, P, P]),
    "esr_alpha2weight_bwd": (I32, [P, P, P, P, P, P, I64, I64, P, P, P, P]),
    "esr_segment_sum_fwd": (I32, [P, P, I64, I32, I64, P, P]),
    "esr_segment_sum_bwd": (I32, [P, P, I64, I32, P, P]),
    "esr_tv_add_grad": (I32, [P, P, F32, F32, F32, I64, I64, I64, I64, I32, P]),
    "esr_exclusive_scan_i32": (I32, [P, P, I64, P, P]),
    "esr_march_count": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P]),
    "esr_march_fill": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P]),
    "esr_march_count_bits": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, I32, P, P]),
    "esr_mask_class_bytes": (I64, [SCENE_P]),
    "esr_mask_classify": (I32, [SCENE_P, P, P, P]),
    "esr_march_fill_bits": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P, I32, P]),
    "esr_alpha_scan_count": (I32, [SCENE_P, P, I64, P, P, P, P, P, P, P]),
    "esr_alpha_scan_fill": (I32, [SCENE_P, P, I64, P, P, P, P, P, P, P, P, P, P, P, P]),
    "esr_alpha_scan_bwd": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_alpha_bwd": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P, I64, P, P]),
    "esr_alpha_scan_count_g": (I32, [SCENE_P, P, I64, P, P, P, P, P, P, P, P]),
    "esr_alpha_scan_bwd_g": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_cos_fwd": (I32, [SCENE_P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_cos_bwd": (I32, [SCENE_P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_cos_vol_fwd": (I32, [SCENE_P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_cos_vol_bwd": (I32, [SCENE_P, P, P, P, P, P, P, I64, P, P]),
    "esr_neus_alpha_bwd_g": (I32, [SCENE_P, P, P, P, I64, P, P, P, P, P, P, P, P, I64, P, P]),
    "esr_encode_coarse_fwd": (I32, [SCENE_P, P, P, P, P, P, P, P, P, I64, P, P]),
    "esr_encode_coarse_bwd": (I32, [SCENE_P, P, P, P, P, P, I64, P, P, P, P, P]),
    "esr_encode_fwd": (I32, [SCENE_P, P, P, P, P, P, P, I32, P, P, P, I64, P, I32, P]),
    "esr_encode_bwd": (I32, [SCENE_P, P, P, P, I32, P, P, I64, P, P, P, P, P]),
    "esr_encode_pbr_fwd": (I32, [SCENE_P, P, P, P, P, P, P, P, I32, P, P, P, P, I64, P, P, I32, P, P]),
    "esr_encode_pbr_bwd": (I32, [SCENE_P, P, P, P, I32, P, P, P, I64, P, P, P, P, P, P, P, P]),
    "esr_sample_points": (I32, [SCENE_P, P, P, P, P, I64, P, P]),
    "esr_sdf_expgrad_fwd": (I32, [SCENE_P, P, P, I64, I32, P, P, P]),
    "esr_sdf_expgrad_bwd": (I32, [SCENE_P, P, I64, P, P, P, P]),
    "esr_lts_accumulate_fwd": (I32, [P, P, P, P, P, P, P, P, P, I64, I32, P, P, P, P, I32, P]),
    "esr_lts_accumulate_bwd": (I32, [P, P, P, P, P, P, P, P, P, I64, I32, P, P, P, P, P, P, P, P, P, I32, P, P]),
    "esr_lts_scatter_dirs": (I32, [P, P, P, I64, I32, P, P]),
    "esr_sg_envmap_fwd": (I32, [P, P, P, P, I32, I32, P, P, I64, P, P]),
    "esr_sg_envmap_bwd": (I32, [P, P, P, P, I32, I32, P, I64, P, P, P, P, P, P]),
    "esr_sdf_fd_gradient": (I32, [SCENE_P, P, P, P, P, P, I64, P, P]),
    "esr_tonemap_encode_fwd": (I32, [P, P, P, P, I64, P, P, I32, P]),
    "esr_tonemap_encode_bwd": (I32, [P, P, P, I64, P, P]),
    "esr_composite_fwd": (I32, [P, I64, P, P, P, P, P, P, P]),
    "esr_composite_bwd": (I32, [P, P, P, P, P, P, P, I64, P, P, P, P]),
    "esr_dvgo_fwd": (I32, [DVGO_P, P, P, P, P, P, P, P, I64, I32, P, P, P, P, P, P, P, P]),
    "esr_dvgo_eval": (I32, [DVGO_P, P, P, P, P, P, I64, I32, P, P, P, P, P, P, P, P, P, P]),
    "esr_dvgo_bwd": (I32, [DVGO_P, P, P, P, P, P, I64, I32, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    "esr_tonemap_mlp_fwd": (I32, [DESC_P, P, P, I64, P, P]),
    "esr_tonemap_mlp_bwd": (I32, [DESC_P, P, P, P, P, P, I64, P, P, P]),
    "esr_adam_step": (I32, [P, P, P, P, P, I64, F32, F32, F32, F32, F32, I64, P]),
    "esr_grad_pack_floats": (I64, [P, I32, I64]),
    "esr_grad_pack": (I32, [P, P, I32, P, I64, P, P]),
    "esr_grad_unpack": (I32, [P, P, I32, P, I64, P, P]),
    "esr_grad_block_flags": (I32, [P, P, I32, I32, I32, I32, I32, I32, I32, P, P]),
    "esr_mlp_param_count": (I64, [DESC_P]),
    "esr_mlp_image_bytes": (I64, [DESC_P]),
    "esr_mlp_act_rows": (I64, [I64]),
    "esr_mlp_hidden_bytes": (I64, [DESC_P, I64]),
    "esr_mlp_dz_bytes": (I64, [DESC_P, I64]),
    "esr_mlp_pack": (I32, [DESC_P, P, P, P]),
    "esr_mlp_fwd": (I32, [DESC_P, P, P, I64, I64, I64, P, P, I64, P]),
    "esr_mlp_bwd": (I32, [DESC_P, P, P, P, P, I64, I64, I64, P, P, P, P, I32, I32, P, P]),
    "esr_mlp_bwd_weights": (I32, [DESC_P, P, I64, I64, I64, P, P, P, P]),
    "esr_render_voxurff_workspace_bytes": (I64, [ctypes.POINTER(Scene), I64, I64, I64, I32]),
    "esr_render_voxurff_fwd": (I32, [ctypes.POINTER(VoxurffStep), P, P, P, P, I64, P, P, P, P]),
    "esr_render_voxurff_bwd": (I32, [ctypes.POINTER(VoxurffStep), P, P, P, P, P, P, P, P, P, P, P, P]),
    "esr_rows_to_mlp_tiles": (I32, [P, I64, I32, P, I32, P, P]),
    "esr_grid_tv_fwd": (I32, [P, P, I32, I64, I64, I64, I64, I64, I64, I64, P, P]),
    "esr_grid_tv_bwd": (I32, [P, P, I32, I64, I64, I64, I64, I64, I64, I64, P, P, F32, P, P]),
    "esr_sdf_central_gradient": (I32, [P, I64, I64, I64, F32, P, P]),
    "esr_smooth_grad_tv_fwd": (I32, [P, P, I64, I64, I64, P, F32, P, P, P]),
    "esr_smooth_grad_tv_bwd": (I32, [P, I64, I64, I64, F32, P, P, F32, P, P]),
}


def lib() -> ctypes.CDLL:
    """Load (building first if the sources are newer) the C-ABI library."""
    global _lib
    if _lib is None:
        path = build()
        try:
            l = ctypes.CDLL(path)
        except OSError as e:  # no silent fallback
            raise EsrError(f"cannot load {path}: {e}") from e
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise EsrError(f"libesr_b200 error {status}: {lib().esr_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EsrError("libesr_b200 operates on CUDA tensors only (no CPU path)")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f3(v) -> F3:
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().tolist()
    return F3(*[float(x) for x in v])
