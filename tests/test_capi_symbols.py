"""The C-ABI library loads and exports every symbol include/speedyml_engine.h declares (no GPU needed)."""
import ctypes
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "speedyml_engine.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sml_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    eng = importlib.import_module("speedy-ml_b200.engine")
    build = importlib.import_module("speedy-ml_b200.build")
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python mirror binds exactly the declared set
    assert sorted(eng.EXPORTED_SYMBOLS) == names


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    eng = importlib.import_module("speedy-ml_b200.engine")
    try:
        eng.Engine(number_of_regions=1152)
    except eng.EngineError as e:
        assert "no CUDA device" in str(e) or "CPU fallback" in str(e)
    else:
        raise AssertionError("engine creation must fail without a CUDA device (no CPU fallback)")


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "speedy-ml_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inl", ".cpp", ".h", ".f90")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle_c", "oracle_np", "speedyml_oracle", "orc_"):
                    assert needle not in src, f"{f} uses the oracle ({needle})"


def test_cpp_host_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    """the compiled host program on speedy-ml_b200/host/speedyml_host.hpp links against the same C ABI (g++ only)"""
    import subprocess

    import numpy as np
    import torch
    build = importlib.import_module("speedy-ml_b200.build")
    exe = build.build_host_driver()
    assert os.access(exe, os.X_OK)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "usage" in p.stderr
    if torch.cuda.is_available():
        return
    case = tmp_path / "empty.bin"
    with open(case, "wb") as f:
        f.write(b"SMLCASE1")
        np.array([1152, 1, 1, 1, 0, 1, 1152, 1, 0], dtype=np.int32).tofile(f)
    p = subprocess.run([exe, str(case), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr      # no CPU fallback in the compiled host either


def test_fortran_shim_binds_only_declared_symbols_with_matching_arity():
    """speedyml_gpu.f90 cannot be compiled here (no Fortran front-end); at least keep it in step with the header:
    every bind(C, name=...) must be a declared entry point with the same number of arguments"""
    hdr = open(os.path.join(ROOT, "include", "speedyml_engine.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(sml_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    f90 = open(os.path.join(ROOT, "speedy-ml_b200", "fortran", "speedyml_gpu.f90")).read()
    f90 = re.sub(r"&\s*\n\s*", " ", f90)
    bound = re.findall(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)", f90)
    assert len(bound) >= 30
    for fname, fargs, cname in bound:
        assert fname == cname
        assert cname in decl, f"{cname} bound in the Fortran layer but not declared in the header"
        n = len([a for a in fargs.split(",") if a.strip()])
        assert n == decl[cname], f"{cname}: {n} Fortran arguments vs {decl[cname]} in the header"
