"""The C-ABI library loads and exports every symbol include/mmlb200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mmlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mml_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from mymedialite_b200 import build, _capi
    build.build()
    return ctypes.CDLL(_capi.SO_PATH)


def test_header_declares_the_path():
    names = declared_functions()
    for needed in ("mml_ratings_create", "mml_ratings_csr", "mml_shuffle_apply", "mml_partition_blocks", "mml_sgd_iterate",
                   "mml_sgd_predict", "mml_sgd_evaluate", "mml_wrmf_iterate", "mml_topn_mf", "mml_ctx_create_dist"):
        assert needed in names


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_bindings_cover_the_header():
    from mymedialite_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_functions()


def test_version_and_error_channel(lib):
    lib.mml_version.restype = ctypes.c_char_p
    assert lib.mml_version().decode().startswith("mmlb200") and "sm_100a" in lib.mml_version().decode()
    lib.mml_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mml_last_error(), bytes)


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the context cannot be created: the product path fails loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    st = lib.mml_ctx_create(1, None, ctypes.byref(h))
    assert st != 0 and not h.value
    lib.mml_last_error.restype = ctypes.c_char_p
    assert lib.mml_last_error()


def test_params_defaults_are_the_reference_defaults():
    """MatrixFactorization.cs:87-96, BiasedMatrixFactorization.cs:85-141."""
    from mymedialite_b200 import engine
    p = engine.default_params()
    assert p.num_factors == 10 and abs(p.learn_rate - 0.01) < 1e-9 and p.decay == 1.0
    assert abs(p.regularization - 0.015) < 1e-9 and abs(p.reg_u - 0.015) < 1e-9 and abs(p.reg_i - 0.015) < 1e-9
    assert p.bias_learn_rate == 1.0 and abs(p.bias_reg - 0.01) < 1e-9
    assert p.frequency_regularization == 0 and p.loss == 0 and p.bold_driver == 0 and p.max_threads == 1


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mymedialite_b200/ may import, load or call it."""
    pkg = os.path.join(ROOT, "mymedialite_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libmmloracle" not in text and "mml_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_csharp_binding_declares_every_entry_point():
    """csharp/NativeMethods.cs (the P/Invoke side a maintainer adds to MyMediaLite.dll) is not compiled here -- no .NET
    toolchain -- so at least its DllImport list is held against the header: same names, same argument counts."""
    cs = open(os.path.join(ROOT, "csharp", "NativeMethods.cs")).read()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mmlb200.h")).read(), flags=re.S)
    imported = {}
    for m in re.finditer(r"\[DllImport\(LIB\)\]\s*(?:public\s+)?static\s+extern\s+\w+\s+(mml_\w+)\s*\((.*?)\)\s*;", cs, flags=re.S):
        args = re.sub(r"\[[^\]]*\]", "", m.group(2))            # drop [Out] / [In, Out] attributes
        imported[m.group(1)] = 0 if not args.strip() else args.count(",") + 1
    declared = {}
    for m in re.finditer(r"\b(mml_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        args = m.group(2).strip()
        declared[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    assert sorted(declared) == declared_functions()
    missing = sorted(set(declared) - set(imported))
    assert not missing, missing
    wrong = {n: (imported[n], declared[n]) for n in declared if imported[n] != declared[n]}
    assert not wrong, wrong


def test_header_is_plain_c(tmp_path):
    """include/mmlb200.h is the FFI contract: it must compile as C99 on its own (no C++ types, no torch, no CUDA headers)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "mmlb200.h"\nint main(void) { mml_mf_params p; (void)p; return MML_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
