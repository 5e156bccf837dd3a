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
