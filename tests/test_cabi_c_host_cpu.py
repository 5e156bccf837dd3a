"""The C ABI from a plain-C host (no GPU, no compute): `include/esr_b200.h` is valid C99 and C++, a C program links
against esr_nerf_b200/libesr_b200.so, the struct layouts the C compiler sees are the ones the Python binding's ctypes
classes assume (a drifted field would silently shift every later pointer of esr_voxurff_step_t), and entry points reject
bad arguments with ESR_ERR_BAD_ARG and a message before touching a device."""
import ctypes
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "esr_b200.h")


def test_header_is_c99_and_cxx():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HDR], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", HDR], check=True)


def test_plain_c_host_links_and_sees_the_bindings_struct_layouts(tmp_path):
    from esr_nerf_b200 import _lib

    so = _lib.build()
    exe = str(tmp_path / "host_check")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "chost", "host_check.c"), "-o", exe, "-L", os.path.dirname(so), "-lesr_b200",
                    "-Wl,-rpath," + os.path.dirname(so)], check=True)
    out = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout)

    assert out["version"] == _lib.lib().esr_version() >= 100
    assert out["ESR_OK"] == 0 and out["ESR_ERR_CAPACITY"] == -3
    # bad arguments: an error code + a message, no device needed (this container has none)
    assert out["rc_null"] == out["rc_neg"] == out["rc_step"] == -1 and out["err_nonempty"] == 1
    assert (out["act_rows_0"], out["act_rows_1"], out["act_rows_129"]) == (0, 128, 256)      # 128-row tiles

    for name in ("Scene", "DvgoScene", "MlpDesc", "VoxurffStep"):
        cls = getattr(_lib, name)
        assert ctypes.sizeof(cls) == out["sizeof"][name], name
        declared = {f[0] for f in cls._fields_}
        for field, off in out[name].items():
            assert field in declared, (name, field)
            assert getattr(cls, field).offset == off, (name, field, getattr(cls, field).offset, off)
    # every field of the one-call step is covered (a new field must be added to the C host check as well)
    assert {f[0] for f in _lib.VoxurffStep._fields_} == set(out["VoxurffStep"])
    assert {f[0] for f in _lib.MlpDesc._fields_} == set(out["MlpDesc"])


def test_ctypes_prototypes_agree_with_the_header_signatures():
    """every entry point: same number of parameters in include/esr_b200.h and in the binding's argtypes, and the same
    type class position by position (pointer / int64 / int32 / float) — a swapped or missing argument in a ctypes
    prototype is silent until the kernel reads garbage"""
    import re

    from esr_nerf_b200 import _lib

    hdr = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)

    def c_class(decl: str) -> str:
        if "*" in decl or "[" in decl or "esr_stream_t" in decl:
            return "ptr"
        for t, k in (("int64_t", "i64"), ("int32_t", "i32"), ("uint8_t", "u8"), ("float", "f32"), ("double", "f64"), ("int", "i32")):
            if re.search(rf"\b{t}\b", decl):
                return k
        raise AssertionError(decl)

    def py_class(t) -> str:
        if t in (ctypes.c_int64,):
            return "i64"
        if t in (ctypes.c_int, ctypes.c_int32):
            return "i32"
        if t is ctypes.c_float:
            return "f32"
        if t is ctypes.c_double:
            return "f64"
        return "ptr"      # c_void_p, c_char_p, POINTER(...)

    seen = 0
    for m in re.finditer(r"\b(esr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        decls = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        _, argtypes = _lib.PROTOTYPES[name]
        assert len(decls) == len(argtypes), (name, len(decls), len(argtypes))
        for i, (d, t) in enumerate(zip(decls, argtypes)):
            assert c_class(d) == py_class(t), (name, i, d, t)
        seen += 1
    assert seen == len(_lib.PROTOTYPES)


def test_training_step_example_compiles_and_links(tmp_path):
    """tests/chost/train_step_example.c — INTEGRATION.md's non-Python training step (esr_render_voxurff_fwd / _bwd, the
    capacity protocol, esr_adam_step) as a complete C99 translation unit: must compile warning-free and resolve every
    symbol against the library (not run: no device here)"""
    from esr_nerf_b200 import _lib

    so = _lib.build()
    obj, lib = str(tmp_path / "tse.o"), str(tmp_path / "libtse.so")
    src = os.path.join(ROOT, "tests", "chost", "train_step_example.c")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-c", src, "-o", obj], check=True)
    subprocess.run(["gcc", "-shared", "-o", lib, obj, "-L", os.path.dirname(so), "-lesr_b200", "-Wl,--no-undefined",
                    "-Wl,-rpath," + os.path.dirname(so)], check=True)
    h = ctypes.CDLL(lib)
    assert hasattr(h, "host_train_step") and hasattr(h, "host_workspace_bytes")


def test_shipped_library_is_current_and_blackwell_native():
    """the in-tree libesr_b200.so (it travels to the GPU box as built here) is newer than every source it is built from,
    is sm_100a code only, and its contraction kernels really are tcgen05 / TMEM / bulk-copy code: UTCHMMA (tcgen05.mma),
    LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk) in the SASS of the MLP chain kernels —
    no mma.sync (HMMA) anywhere (scripts/sass_summary.py writes the per-kernel table under profiles/)"""
    import re
    import shutil

    import pytest

    from esr_nerf_b200 import _lib

    so = _lib.build()
    newest = max(os.path.getmtime(os.path.join(_lib.CSRC, f)) for f in list(_lib.SOURCES) + list(_lib.HEADERS))
    assert os.path.getmtime(so) >= newest
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = set()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per_kernel[cur].add(m.group(1))
    assert len(per_kernel) > 90
    assert not any("HMMA" in ops and "UTCHMMA" not in ops for ops in per_kernel.values())      # no mma.sync kernels
    assert not any(op in ("HMMA", "IMMA", "HGMMA") for ops in per_kernel.values() for op in ops)
    chains = {k: ops for k, ops in per_kernel.items() if re.search(r"k_mlp_(fwd|dgrad|wgrad)|k_tonemap_(fwd2|bwd_fused)", k)}
    assert len(chains) >= 30
    for k, ops in chains.items():
        assert {"UTCHMMA", "LDTM", "UTCBAR"} <= ops, (k, sorted(ops & {"UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP"}))
    assert any("UBLKCP" in ops for k, ops in chains.items() if "wgrad" in k)
    assert any("STTM" in ops for k, ops in chains.items() if "k_mlp_fwd" in k)
