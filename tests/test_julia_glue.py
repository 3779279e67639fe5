"""julia/TNB200.jl cannot be executed here (no Julia in the image), so its `ccall`s are checked STATICALLY against the
C ABI: every symbol include/tnb200.h declares has a binding, and each binding's return type, argument count and
argument classes (pointer / 32-bit int / 64-bit int / double / size_t) agree with the ctypes table that the GPU tests
exercise (itensorsgpu.jl_b200/_lib.py: SIGNATURES, itself checked against the header and the exports by
tests/test_abi.py).  A stub that drifts from the header is caught on the CPU."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s):
    """split on commas that are not nested in () or {}"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, i):
    """s[i] == '(' -> index just past its matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced")


def _ccalls():
    src = open(os.path.join(ROOT, "julia", "TNB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)                  # strip comments
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = _balanced(src, m.end() - 1)
        parts = _split_top(src[m.end():end - 1])
        name = re.match(r"\(:(\w+),\s*LIB\)", parts[0]).group(1)
        ret = parts[1]
        assert parts[2].startswith("(") and parts[2].endswith(")"), (name, parts[2])
        types = _split_top(parts[2][1:-1])
        args = parts[3:]
        calls.append((name, ret, types, args))
    return calls


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Cuint": "i32", "Int64": "i64", "UInt64": "i64", "Clonglong": "i64",
            "Float64": "f64", "Cdouble": "f64", "Csize_t": "size"}[t]


def _ct_class(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_int))) and issubclass(t, C._Pointer):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "i64", C.c_double: "f64",
            C.c_size_t: "size"}[t]


def _signatures():
    import sys
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("tnb_lib_table", os.path.join(ROOT, "itensorsgpu.jl_b200", "_lib.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.SIGNATURES


def test_every_declared_entry_point_has_a_julia_binding():
    names = {c[0] for c in _ccalls()}
    sig = _signatures()
    assert sorted(set(sig) - names) == []
    assert sorted(names - set(sig)) == []               # and nothing is bound that the header does not declare


def test_julia_ccall_signatures_match_the_c_abi():
    sig = _signatures()
    size_is_64 = C.sizeof(C.c_size_t) == 8
    norm = (lambda c: "i64" if (c == "size" and size_is_64) else c)
    checked = 0
    for name, ret, types, args in _ccalls():
        res, argtypes = sig[name]
        assert len(types) == len(argtypes), "%s: %d Julia argument types, C ABI has %d" % (name, len(types), len(argtypes))
        assert len(args) == len(types), "%s: %d values passed for %d declared types" % (name, len(args), len(types))
        jl = [_jl_class(t) for t in types]
        ct = [_ct_class(t) for t in argtypes]
        for i, (a, b) in enumerate(zip(jl, ct)):
            # size_t and uint64_t are one ctypes class on LP64 (the only ABI this library targets)
            assert norm(a) == norm(b), "%s: argument %d is %s in Julia (%s) but %s in the C ABI" % (name, i, a, types[i], b)
        assert norm(_jl_class(ret)) == norm(_ct_class(res)), "%s: return type %s vs %s" % (name, ret, res)
        checked += 1
    assert checked >= 50


def test_bond_dims_struct_layout_matches():
    """struct BondDims in Julia mirrors tnb_bond_dims field for field (two Int64 then five Int32)"""
    src = open(os.path.join(ROOT, "julia", "TNB200.jl")).read()
    m = re.search(r"struct BondDims\s+(.*?)\s+end", src, re.S)
    fields = [f.strip() for f in m.group(1).replace("\n", ";").split(";") if f.strip()]
    assert fields == ["chiL::Int64", "chiR::Int64", "d1::Int32", "d2::Int32", "wL::Int32", "wM::Int32", "wR::Int32"]
    hdr = open(os.path.join(ROOT, "include", "tnb200.h")).read()
    body = re.search(r"typedef struct tnb_bond_dims_s \{(.*?)\} tnb_bond_dims;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = [d.strip() for d in body.split(";") if d.strip()]
    assert decl == ["int64_t chiL, chiR", "int32_t d1, d2", "int32_t wL, wM, wR"]
    # tnb_plan_desc: 3 x int64, 4 x int32, 9 arrays of 12 int64, 8 x int32, int64, double -- same in Julia's PlanDesc
    pd = re.search(r"struct PlanDesc\s+(.*?)\nend", src, re.S).group(1)
    jl_fields = [f.strip() for f in pd.replace("\n", ";").split(";") if f.strip()]
    kinds = [f.split("::")[1] for f in jl_fields]
    assert kinds == ["Int64"] * 3 + ["Int32"] * 4 + ["NTuple{12,Int64}"] * 9 + ["Int32"] * 8 + ["Int64", "Float64"]
