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


def test_header_is_plain_c_and_links(fts, tmp_path):
    """include/fts_b200.h is the boundary a foreign host binds: it must compile as strict C99 (no C++-isms, no torch or
    CUDA types) and a C program must link against the library through it."""
    import subprocess
    from fts_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text('#include "fts_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    double x[4], w[4], h; char msg[256];\n'
                   '    if (fts_abi_version() != FTS_ABI_VERSION) return 1;\n'
                   '    if (fts_tanhsinh(FTS_F64, 8, 1, x, w, &h) != FTS_OK) return 2;   /* host-side table generation needs no GPU */\n'
                   '    fts_last_error(msg, sizeof msg);\n'
                   '    printf("%.17g %.17g %.17g\\n", x[0], w[0], h);\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lfts_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    xs, ws, h = fts.tanhsinh(np.float64, 8)
    assert [float(v) for v in out] == [float(xs[0]), float(ws[0]), float(h)]


# ------------------------------------------------------------------ the Julia shim (static check) ---
JULIA_DIR = os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200", "julia", "FastTanhSinhQuadratureB200")


def _split_top(text):
    """split on commas that are outside (), {} and []"""
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _julia_ccalls(src):
    """-> [(symbol, number of argument types, number of arguments passed)] of every ccall((:sym, LIB[]), ret, (types...), args...)"""
    out = []
    for m in re.finditer(r"ccall\(\(:(fts_\w+),\s*LIB(?:\[\])?\),", src):
        i, depth = m.end(), 1                     # inside ccall( ... ): find its closing parenthesis
        j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        parts = _split_top(src[i:j - 1])         # ret, (types...), args...
        types = parts[1].strip()
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        tl = _split_top(types[1:-1])
        out.append((m.group(1), len(tl), len(parts) - 2, tl, parts[0].strip()))
    return out


def _c_class(param):
    """width class of a C parameter declaration of the header"""
    t = param.strip()
    if "*" in t or re.search(r"\bfts_(cache|integrand|tables)\b", t):
        return "ptr"
    t = re.sub(r"\s*\b\w+$", "", t).replace("const", "").strip()  # drop the parameter name
    return {"int": "i32", "int32_t": "i32", "fts_dtype": "i32", "int64_t": "i64", "uint64_t": "i64", "long long": "i64", "size_t": "usize",
            "double": "f64", "float": "f32", "unsigned": "i32", "unsigned int": "i32"}[t]


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Cuint": "i32", "Int64": "i64", "UInt64": "i64", "Csize_t": "usize", "Cdouble": "f64", "Cfloat": "f32"}[t]


def test_julia_shim_binds_the_declared_entry_points(fts):
    """Julia is not in the image, so the shim cannot be executed here: check statically that every `ccall` names an
    exported entry point of include/fts_b200.h with the header's parameter count (and passes as many arguments as it
    declares types), that the shim exports the reference's 21 names, and that its family ids are the header's."""
    from fts_b200 import _lib
    src = open(os.path.join(JULIA_DIR, "src", "FastTanhSinhQuadratureB200.jl")).read()
    protos = {}
    for ret, name, params in re.findall(r"^\s*(int|size_t|int64_t)\s+(fts_\w+)\s*\(([^;]*?)\)\s*;", HEADER, flags=re.M | re.S):
        params = re.sub(r"/\*.*?\*/", "", params, flags=re.S).strip()
        protos[name] = (ret, [] if params in ("", "void") else [_c_class(q) for q in _split_top(params)])
    calls = _julia_ccalls(src)
    assert len(calls) >= 20
    doc_calls = _julia_ccalls(open(os.path.join(ROOT, "INTEGRATION.md")).read())  # the bindings shown to a maintainer
    assert len(doc_calls) >= 6
    calls = calls + doc_calls
    lib = C.CDLL(_lib.LIB_PATH)
    for sym, ntypes, nargs, types, ret in calls:
        assert sym in protos and hasattr(lib, sym), f"{sym}: not an exported entry point"
        cret, cparams = protos[sym]
        assert ntypes == len(cparams), f"{sym}: ccall declares {ntypes} argument types, the header {len(cparams)} parameters"
        assert nargs == ntypes, f"{sym}: {nargs} arguments passed for {ntypes} declared types"
        assert [_jl_class(t) for t in types] == cparams, f"{sym}: {types} against {cparams}"
        assert _jl_class(ret) == {"int": "i32", "size_t": "usize", "int64_t": "i64"}[cret], (sym, ret, cret)
    exported = set(re.findall(r"\b\w+\b", re.search(r"^export (.*?)\n\n", src, flags=re.M | re.S).group(1)))
    assert set(fts.REFERENCE_EXPORTS) <= exported, set(fts.REFERENCE_EXPORTS) - exported
    ids = dict((k, int(v)) for k, v in re.findall(r":(\w+) => (\d+)", re.search(r"const FAMILY_IDS = Dict\((.*?)\)\n", src, flags=re.S).group(1)))
    assert ids == dict(fts.FAMILIES), set(ids.items()) ^ set(dict(fts.FAMILIES).items())
    enums = dict((k, int(v, 0)) for k, v in re.findall(r"\b(FTS_[A-Z0-9_]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", HEADER))
    jl = dict(re.findall(r"const ((?:\w+, )*\w+) = ((?:\w+\(\d+\), )*\w+\(\d+\))", src))
    consts = {}
    for names, vals in jl.items():
        for n, v in zip(names.split(", "), re.findall(r"\((\d+)\)", vals)):
            consts[n] = int(v)
    for jn, hn in (("FTS_F32", "FTS_F32"), ("FTS_F64", "FTS_F64"), ("FTS_DD64", "FTS_DD64"), ("ORDER_REFERENCE", "FTS_ORDER_REFERENCE"), ("ORDER_FAST", "FTS_ORDER_FAST"),
                   ("OPT_QUAD_EDGE_RULES", "FTS_OPT_QUAD_EDGE_RULES"), ("OPT_FORCE_DEPTH", "FTS_OPT_FORCE_DEPTH"), ("FLAG_CONVERGED", "FTS_FLAG_CONVERGED"),
                   ("FLAG_MAX_LEVELS", "FTS_FLAG_MAX_LEVELS"), ("KIND_1D", "FTS_KIND_1D"), ("KIND_2D", "FTS_KIND_2D"), ("KIND_3D", "FTS_KIND_3D"),
                   ("KIND_1D_CMPL", "FTS_KIND_1D_CMPL"), ("KIND_2D_CMPL", "FTS_KIND_2D_CMPL")):
        assert consts[jn] == enums[hn], (jn, consts.get(jn), enums.get(hn))
    # the Integrals.jl extension is registered and only calls names the shim defines
    toml = open(os.path.join(JULIA_DIR, "Project.toml")).read()
    assert "FastTanhSinhQuadratureB200IntegralsExt" in toml and os.path.exists(os.path.join(JULIA_DIR, "ext", "FastTanhSinhQuadratureB200IntegralsExt.jl"))
    ext = open(os.path.join(JULIA_DIR, "ext", "FastTanhSinhQuadratureB200IntegralsExt.jl")).read()
    for name in set(re.findall(r"FTS\.(\w+)", ext)):
        assert re.search(rf"^(?:function |struct |const )?{name}\b", src, flags=re.M), name
    for a, b in ("()", "[]", "{}"):
        assert src.count(a) == src.count(b) and ext.count(a) == ext.count(b)
