"""Static consistency of the Julia extension (julia/ext/B200Ext.jl) with include/insb200.h.  The image has no Julia,
so the binding cannot be executed here; what can be checked without it is that the two C structs are mirrored field
by field and that every `ccall` names a declared symbol with the right number and kind of arguments."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "insb200.h")
JULIA = os.path.join(ROOT, "julia", "ext", "B200Ext.jl")


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def c_struct_fields(name):
    """[(field, kind, count)] in declaration order; kind in {int32, int64, double, ptr}."""
    src = _strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const void \*|long long |unsigned long long |double |int )(.*)", decl)
        assert m, decl
        base = m.group(1).strip()
        kind = {"const void *": "ptr", "long long": "int64", "unsigned long long": "int64", "double": "double", "int": "int32"}[base]
        for d in _split_top(m.group(2)):
            dm = re.match(r"\*?\s*(\w+)((?:\[\d+\])*)$", d)
            assert dm, d
            count = 1
            for n in re.findall(r"\[(\d+)\]", dm.group(2)):
                count *= int(n)
            fields.append((dm.group(1), kind, count))
    return fields


JL_KIND = {"Cint": "int32", "Clonglong": "int64", "Culonglong": "int64", "Cdouble": "double", "Ptr{Cvoid}": "ptr"}


def julia_struct_fields(name):
    src = open(JULIA).read()
    body = re.search(r"^struct %s\n(.*?)^end" % name, src, flags=re.S | re.M).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        fname, ftype = [t.strip() for t in line.split("::")]
        m = re.match(r"NTuple\{(\d+),\s*(.+)\}$", ftype)
        count, base = (int(m.group(1)), m.group(2)) if m else (1, ftype)
        fields.append((fname, JL_KIND[base], count))
    return fields


def test_julia_structs_mirror_the_header_field_by_field():
    for cname, jname in (("insb200_setup", "CSetup"), ("insb200_params", "CParams")):
        c, j = c_struct_fields(cname), julia_struct_fields(jname)
        assert [f[0] for f in c] == [f[0] for f in j], (cname, [f[0] for f in c], [f[0] for f in j])
        assert c == j, [(a, b) for a, b in zip(c, j) if a != b]


def c_prototypes():
    """{symbol: [kind per parameter]}; kind in {int32, int64, double, ptr}."""
    src = _strip_c_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*|long long|double)\s*\**\s*(insb200_[a-z0-9_]+)\s*\((.*?)\);", src, flags=re.S):
        args = " ".join(m.group(2).split())
        kinds = []
        if args != "void":
            for a in _split_top(args):
                if "*" in a or "(" in a or "[" in a or a.startswith("insb200_step_callback_t"):  # pointers, call-backs, arrays
                    kinds.append("ptr")
                elif re.match(r"(unsigned )?long long\b", a):
                    kinds.append("int64")
                elif a.startswith("double"):
                    kinds.append("double")
                else:
                    assert re.match(r"(unsigned )?int\b", a), a
                    kinds.append("int32")
        protos[m.group(1)] = kinds
    return protos


def julia_ccalls():
    """[(symbol, return type, [argument types])] of every ccall into libinsb200."""
    src = open(JULIA).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), libinsb200\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = [t for t in _split_top(src[m.end():i - 1].replace("\n", " ")) if t]
        calls.append((m.group(1), m.group(2), types))
    return calls


def _jl_kind(t):
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return JL_KIND[t]


def test_every_ccall_matches_a_declared_prototype():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 18
    for sym, ret, types in calls:
        assert sym in protos, f"{sym} is called from Julia but not declared in insb200.h"
        want = protos[sym]
        got = [_jl_kind(t) for t in types]
        assert got == want, (sym, got, want)
        assert ret in ("Cint", "Cstring", "Clonglong", "Cdouble"), (sym, ret)


def test_entry_points_the_integration_table_promises_are_bound():
    bound = {c[0] for c in julia_ccalls()}
    for sym in ("insb200_plan_create", "insb200_plan_destroy", "insb200_psolve", "insb200_project", "insb200_step_rk",
                "insb200_step_lmwray3", "insb200_rhs", "insb200_stage_combine_project", "insb200_axpy", "insb200_cfl",
                "insb200_solve_unsteady", "insb200_random_field", "insb200_right_hand_side", "insb200_set_bc_u_plane",
                "insb200_apply_bc_u", "insb200_energy_spectrum", "insb200_spectral_moments"):
        assert sym in bound, sym


def test_python_ctypes_signatures_match_the_header_too():
    """The same check for the harness the tests drive: insb200/_lib.py's (restype, argtypes) table against the
    prototypes of include/insb200.h, parameter by parameter."""
    import ctypes as C

    from insb200 import _lib

    def kind(t):
        if t in (C.c_int, C.c_uint):
            return "int32"
        if t in (C.c_longlong, C.c_ulonglong):
            return "int64"
        if t is C.c_double:
            return "double"
        return "ptr"  # c_void_p, POINTER(...), CFUNCTYPE call-backs, c_char_p

    protos = c_prototypes()
    assert set(protos) == set(_lib._PROTOTYPES)
    for sym, (res, args) in _lib._PROTOTYPES.items():
        assert [kind(a) for a in args] == protos[sym], (sym, [kind(a) for a in args], protos[sym])
