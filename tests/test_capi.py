"""The C-ABI library loads on a CPU-only host, exports every symbol include/dune_copasi_b200.h
declares, and refuses to compute without a CUDA device (no silent CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cases as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dune_copasi_b200.h")
LIB = os.path.join(ROOT, "dune_copasi_b200", "libdune_copasi_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from dune_copasi_b200 import capi
    names = declared_symbols()
    assert len(names) >= 60
    lib = ctypes.CDLL(LIB)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(capi.SYMBOLS) == names, set(capi.SYMBOLS) ^ set(names)
    exported = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exp = sorted(set(re.findall(r" T (dcb_[a-z0-9_]+)", exported)))
    assert exp == names, set(exp) ^ set(names)


def test_no_torch_types_and_extern_c():
    text = open(HEADER).read()
    assert 'extern "C"' in text and "torch" not in text.lower() and "std::" not in text


def test_config_roundtrip_and_overrides():
    import dune_copasi_b200 as D
    cfg = D.Config("[a.b]\nc = 1 # comment\nd.e = x y z\n[f]\ng=2\n")
    cfg.set("a.b.c", "3").set("new.key", "v")
    dump = cfg.dump()
    assert "a.b.c = 3" in dump and "a.b.d.e = x y z" in dump and "f.g = 2" in dump and "new.key = v" in dump


@pytest.mark.skipif(__import__("dune_copasi_b200").lib().dcb_device_count() > 0, reason="needs a host without GPU")
def test_compute_entry_points_fail_loudly_without_gpu():
    import dune_copasi_b200 as D
    cfg, model, grid = K.product_objects(K.CASES["exp"])
    with pytest.raises(D.DcbError, match="no CUDA device"):
        D.Operator(model, grid)


def test_library_does_not_link_the_oracle():
    deps = subprocess.check_output(["ldd", LIB], text=True)
    assert "oracle" not in deps
    syms = subprocess.check_output(["nm", "-D", LIB], text=True)
    assert "orc_" not in syms
    # the product package never imports the oracle
    for dp, _, fs in os.walk(os.path.join(ROOT, "dune_copasi_b200")):
        for f in fs:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def _build_c_host(tmp_path):
    """tests/c_host/host_check.c: a plain C99 program on the C ABI (no Python, no torch)."""
    exe = str(tmp_path / "host_check")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_host", "host_check.c"), "-o", exe,
                           "-L", os.path.dirname(LIB), "-ldune_copasi_b200", "-Wl,-rpath," + os.path.dirname(LIB)])
    return exe


def test_header_is_c99_and_a_c_host_runs(tmp_path):
    import dune_copasi_b200 as D
    exe = _build_c_host(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    f = out.stdout.split()
    # 5 x 4 x 3 vertices, 2 species: 120 rows; 15-point (Kuhn P1) and 27-point (Q1) stencils
    assert f[0] == "ok" and f[1] == "120" and f[3] == "120" and int(f[2]) < int(f[4])
    if D.lib().dcb_device_count() < 1:
        assert f[5] == "nodevice"         # no silent CPU fallback behind the C ABI either


@pytest.mark.gpu
def test_c_host_computes_on_the_device(tmp_path):
    exe = _build_c_host(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[0] == "ok" and float(out.stdout.split()[5]) > 0


def test_peer_mailbox_layout(tmp_path):
    """csrc/kernels/peer.hpp: the slots of the CUDA-IPC mailboxes never overlap (host-side check)."""
    exe = str(tmp_path / "peer_layout")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "dune_copasi_b200", "csrc"),
                           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "tests", "c_host", "peer_layout_check.cpp"),
                           "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def _build_cxx_example(tmp_path):
    exe = str(tmp_path / "gray_scott")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "gray_scott.cpp"), "-o", exe,
                           "-L", os.path.dirname(LIB), "-ldune_copasi_b200", "-Wl,-rpath," + os.path.dirname(LIB)])
    return exe


def test_cxx_example_builds_and_refuses_to_run_without_a_device(tmp_path):
    """examples/gray_scott.cpp: the documented Gray-Scott model end to end from C++ (no Python)."""
    import dune_copasi_b200 as D
    exe = _build_cxx_example(tmp_path)
    out = subprocess.run([exe, "16", "1.0"], capture_output=True, text=True, timeout=300)
    if D.lib().dcb_device_count() < 1:
        assert out.returncode == 2 and "no CUDA device" in out.stderr
    else:
        assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_cxx_example_runs_on_the_device(tmp_path):
    exe = _build_cxx_example(tmp_path)
    out = subprocess.run([exe, "64", "5.0", str(tmp_path / "vtk")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("t = 5 after") and (tmp_path / "vtk" / "vtk-compartment-00000.vtu").exists()


def _build_dune_shim(tmp_path):
    """examples/dune_shim: the reference-side binding of INTEGRATION.md section 2 (B200StageOperator, B200LinearSolver)
    compiled against stand-in PDELab names and linked with the C ABI"""
    exe = str(tmp_path / "shim_check")
    shim = os.path.join(ROOT, "examples", "dune_shim")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(shim, "stub"), "-I", shim, os.path.join(shim, "check.cpp"), "-o", exe,
                           "-L", os.path.dirname(LIB), "-ldune_copasi_b200", "-Wl,-rpath," + os.path.dirname(LIB)])
    return exe


def test_dune_side_shim_compiles_and_fails_loudly_without_a_device(tmp_path):
    import dune_copasi_b200 as D
    out = subprocess.run([_build_dune_shim(tmp_path)], capture_output=True, text=True, timeout=300)
    if D.lib().dcb_device_count() < 1:
        assert out.returncode == 2 and "no CUDA device" in out.stderr
    else:
        assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_dune_side_shim_runs_on_the_device(tmp_path):
    """residual through the shim == residual through the ABI, and J z = b solved back to z through B200LinearSolver"""
    out = subprocess.run([_build_dune_shim(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("shim ok"), out.stdout + out.stderr
