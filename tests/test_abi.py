"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/fts_b200.h declares, agrees with the header's enums, and refuses to compute
without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "fts_b200.h")).read()


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_library_exports_every_declared_symbol(fts):
    from fts_b200 import _lib
    declared = sorted(set(re.findall(r"^\s*(?:int|size_t|int64_t)\s+(fts_\w+)\s*\(", HEADER, flags=re.M)))
    assert len(declared) >= 30
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, f"libfts_b200.so does not export {missing}"
    # the Python binding declares the same set
    assert sorted(_lib.EXPORTED) == declared
    assert not _lib.MISSING
    assert lib.fts_abi_version() == int(re.search(r"#define FTS_ABI_VERSION (\d+)", HEADER).group(1))


def test_enums_match_header(fts):
    from fts_b200 import _lib
    enums = dict((k, int(v, 0)) for k, v in re.findall(r"\b(FTS_[A-Z0-9_]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", HEADER))
    assert enums["FTS_OK"] == _lib.OK and enums["FTS_ERR_NO_DEVICE"] == _lib.ERR_NO_DEVICE
    assert (enums["FTS_F32"], enums["FTS_F64"], enums["FTS_DD64"]) == (_lib.F32, _lib.F64, _lib.DD64)
    assert (enums["FTS_ORDER_REFERENCE"], enums["FTS_ORDER_FAST"], enums["FTS_ORDER_AUTO"]) == (0, 1, 2)
    assert enums["FTS_FLAG_CONVERGED"] == fts.FLAG_CONVERGED and enums["FTS_FLAG_MAX_LEVELS"] == fts.FLAG_MAX_LEVELS
    assert enums["FTS_KIND_1D_CMPL"] == fts.KIND_1D_CMPL and enums["FTS_KIND_2D_CMPL"] == fts.KIND_2D_CMPL
    # builtin family ids: header == oracle table == python table == library's own info
    from fts_oracle import FAMILIES as ORACLE_FAMILIES
    header_fams = {k[4:]: v for k, v in enums.items() if re.match(r"FTS_[FC][123]_", k)}
    assert header_fams == ORACLE_FAMILIES
    assert sorted(fts.FAMILIES.values()) == sorted(header_fams.values())
    for name, fid in fts.FAMILIES.items():
        f = fts.builtin(name)
        kind, npar = C.c_int(), C.c_int()
        assert _lib.lib.fts_family_info(fid, C.byref(kind), C.byref(npar)) == 0
        assert (f.kind, f.nparams) == (kind.value, npar.value)


def test_reference_export_surface(fts):
    """the 21 exported names of src/FastTanhSinhQuadrature.jl:44-51 exist with the same defaults"""
    import inspect
    for name in fts.REFERENCE_EXPORTS:
        assert callable(getattr(fts, name)), name
    assert len(fts.REFERENCE_EXPORTS) == 21
    d = lambda fn, kw: inspect.signature(fn).parameters[kw].default
    assert d(fts.adaptive_integrate_1D, "max_levels") == 16 and d(fts.adaptive_integrate_2D, "max_levels") == 8
    assert d(fts.adaptive_integrate_3D, "max_levels") == 5 and d(fts.adaptive_integrate_1D_cmpl, "max_levels") == 16
    assert d(fts.adaptive_integrate_1D, "rtol") is None and d(fts.adaptive_integrate_1D, "atol") == 0
    assert d(fts.adaptive_integrate_1D, "warn") is True and d(fts.adaptive_integrate_1D, "cache") is None
    assert d(fts.adaptive_cache_1D, "max_levels") == 16 and d(fts.adaptive_cache_2D, "max_levels") == 8
    assert d(fts.adaptive_cache_3D, "max_levels") == 5 and d(fts.adaptive_cache_1D, "complement") is False


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_compute_fails_loudly_without_a_device(fts):
    x, w, h = fts.tanhsinh(np.float64, 16)
    with pytest.raises(fts.FtsError) as ei:
        fts.integrate1D(fts.builtin("exp"), 0.0, 1.0, x, w, h)
    assert ei.value.code == 6 and "no CPU fallback" in str(ei.value)
    with pytest.raises(fts.FtsError):
        fts.quad(fts.builtin("exp"), 0.0, 1.0)


def test_product_never_touches_the_oracle():
    """the shipped package and library may not import, link or mention the test oracle"""
    pkg = os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200")
    offenders = []
    for base, _dirs, files in os.walk(pkg):
        if os.path.basename(base) in ("build", "lib", "__pycache__"):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")) or fn == "Makefile":
                txt = open(os.path.join(base, fn), errors="replace").read()
                if re.search(r"fts_oracle|fts_or_|oracle/", txt):
                    offenders.append(os.path.join(base, fn))
    assert not offenders, offenders
    out = os.popen(f"ldd {os.path.join(pkg, 'lib', 'libfts_b200.so')}").read()
    assert "oracle" not in out
