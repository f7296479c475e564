"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/rass_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rassengine_b200 import build, _capi
    build.build()
    return _capi.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "rass_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rass_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from rassengine_b200 import _capi
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _capi.PROTOTYPES, f"{n} has no ctypes prototype"


def test_version_and_no_cpu_fallback(lib):
    import torch
    assert b"sm_100a" in lib.rass_version()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.rass_create(1024, 0, 0, 0, 0, ctypes.byref(h))
    assert rc == -3 and not h.value                    # RASS_E_CUDA
    assert b"no CPU fallback" in lib.rass_last_error(None)
    import rassengine_b200 as rb
    with pytest.raises(rb.RassError):
        rb.Engine(1024)


def test_invalid_arguments_are_rejected_before_cuda(lib):
    h = ctypes.c_void_p()
    assert lib.rass_create(0, 0, 0, 0, 0, ctypes.byref(h)) == -1
    assert lib.rass_create(2048, 0, 0, 0, 0, ctypes.byref(h)) == -1
    assert lib.rass_create(1024, 7, 0, 0, 0, ctypes.byref(h)) == -1
    assert lib.rass_create(1024, 0, 0, 0, 3, ctypes.byref(h)) == -1
    assert lib.rass_count(None, None) == -1


def _declarations():
    """name -> list of parameter declarations, parsed from the header (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "rass_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"\b(int|const char\s*\*)\s*(rass_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = " ".join(params.split())
        out[name] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return out


def test_ctypes_prototypes_follow_the_header():
    """The Python binding (what INTEGRATION.md shows a maintainer) must not drift from include/rass_b200.h: same
    arity, pointers bound as pointers, 64-bit integers as 64-bit, the status / string return types."""
    import ctypes as C
    from rassengine_b200 import _capi
    decl = _declarations()
    assert set(decl) == set(_capi.PROTOTYPES), set(decl) ^ set(_capi.PROTOTYPES)
    for name, params in decl.items():
        res, args = _capi.PROTOTYPES[name]
        assert len(params) == len(args), (name, params, args)
        for p, a in zip(params, args):
            is_ptr = "*" in p
            bound_ptr = a in (C.c_void_p, C.c_char_p) or (isinstance(a, type) and issubclass(a, C._Pointer))
            assert is_ptr == bound_ptr, (name, p, a)
            if not is_ptr:
                base = p.rsplit(" ", 1)[0].replace("const ", "").strip()
                want = {"int": C.c_int, "int64_t": C.c_int64, "uint32_t": C.c_uint32, "float": C.c_float,
                        "double": C.c_double, "size_t": C.c_size_t}[base]
                assert a is want, (name, p, a)
        assert res is (C.c_char_p if name in ("rass_version", "rass_last_error") else C.c_int), name


def test_stats_struct_and_constants_follow_the_header():
    import ctypes as C
    from rassengine_b200 import _capi
    text = open(os.path.join(ROOT, "include", "rass_b200.h")).read()
    body = re.search(r"typedef struct rass_stats \{(.*?)\} rass_stats;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [tuple(f.split()) for f in body.split(";") if f.strip()]
    ctype = {"double": C.c_double, "int64_t": C.c_int64, "int32_t": C.c_int32}
    assert [(n, ctype[t]) for t, n in fields] == list(_capi.RassStats._fields_)
    # enumerators / defines the binding mirrors as module constants
    consts = dict(re.findall(r"\b(RASS_[A-Z0-9_]+)\s*=\s*(-?\d+)", text))
    consts.update(dict(re.findall(r"#define\s+(RASS_[A-Z0-9_]+)\s+(-?\d+)\b", text)))
    mirrored = {"RASS_PATH_AUTO": "PATH_AUTO", "RASS_PATH_STREAM": "PATH_STREAM", "RASS_PATH_UMMA": "PATH_UMMA",
                "RASS_PATH_EXACT": "PATH_EXACT", "RASS_PATH_GEMM": "PATH_GEMM", "RASS_METRIC_COSINE": "METRIC_COSINE",
                "RASS_METRIC_L2": "METRIC_L2", "RASS_OPT_KNN_PREFILTER": "OPT_KNN_PREFILTER", "RASS_E_AGAIN": "RASS_E_AGAIN"}
    for c_name, py_name in mirrored.items():
        assert c_name in consts, c_name
        assert int(consts[c_name]) == getattr(_capi, py_name), c_name
