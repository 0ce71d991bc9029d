"""Static cross-check of the Julia binding (easysimp.jl_b200/julia/EasySIMPB200.jl) against the C ABI it binds
(include/easysimp_b200.h).  The build image has no Julia toolchain, so the binding is never executed here; what CAN be held
without running it: every `ccall` names a function the header declares (and the built library exports), passes the right
number of arguments, and every Julia argument / return type is the ccall spelling of the C type in that position; the two
`struct`s that cross the boundary by reference have the header's fields in the header's order with layout-compatible types;
and the entry points of the reference that the binding claims to provide are defined in it."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "easysimp_b200.h"
JULIA = ROOT / "easysimp.jl_b200" / "julia" / "EasySIMPB200.jl"

# C parameter type (const dropped, spaces removed) -> the Julia ccall types that may stand for it
C_TO_JULIA = {
    "int": {"Cint"},
    "int64_t": {"Int64"},
    "double": {"Cdouble"},
    "double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "void*": {"Ptr{Cvoid}", "Ptr{UInt8}"},
    "char*": {"Cstring"},
    "esimp_ctx*": {"Ptr{Cvoid}"},
    "esimp_ctx**": {"Ref{Ptr{Cvoid}}"},
    "esimp_multi*": {"Ptr{Cvoid}"},
    "esimp_multi**": {"Ref{Ptr{Cvoid}}"},
    "esimp_params*": {"Ref{EsimpParams}"},
    "esimp_iter_stats*": {"Ref{EsimpIterStats}"},
}
C_FIELD_TO_JULIA = {"double": {"Cdouble", "Float64"}, "int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}}


def _strip_comments(text):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))


def header_prototypes():
    text = _strip_comments(HEADER.read_text())
    protos = {}
    for m in re.finditer(r"\b((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*\b(esimp_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\bconst\b", "", a).strip()
                mm = re.match(r"([A-Za-z_0-9]+)\s*(\**)\s*[A-Za-z_0-9]*$", a)  # type, stars, optional name
                assert mm, (name, a)
                params.append(mm.group(1) + mm.group(2))
        protos[name] = (re.sub(r"\s+|\bconst\b", "", ret), params)
    return protos


def header_struct(name):
    text = _strip_comments(HEADER.read_text())
    body = re.search(r"typedef\s+struct\s+" + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            ctype, names = decl.split(None, 1)
            fields += [(n.strip(), ctype) for n in names.split(",")]
    return fields


def _split_top(s):
    """split on commas that are not nested in (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(text, start):
    """text[start] == '(' -> index one past its closing parenthesis (string literals skipped)"""
    depth, i, in_str = 0, start, False
    while i < len(text):
        ch = text[i]
        if in_str:
            if ch == "\\":
                i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise AssertionError("unbalanced ccall")


def julia_ccalls():
    text = "\n".join("" if l.lstrip().startswith("#") else l for l in JULIA.read_text().split("\n"))  # comment lines blanked: line numbers hold
    calls = []
    for m in re.finditer(r"\bccall\(", text):
        end = _balanced(text, m.end() - 1)
        parts = _split_top(text[m.end():end - 1])
        sym = re.match(r"\(\s*:(\w+)\s*,\s*LIB\s*\)$", parts[0])
        assert sym, parts[0]
        types = parts[2].strip()
        assert types.startswith("(") and types.endswith(")"), types
        calls.append((sym.group(1), parts[1], _split_top(types[1:-1]), parts[3:], text.count("\n", 0, m.start()) + 1))
    return calls


def julia_struct(name):
    text = JULIA.read_text()
    body = re.search(r"^struct\s+" + name + r"\s*\n(.*?)^end", text, flags=re.S | re.M).group(1)
    return [(f.split("::")[0].strip(), f.split("::")[1].strip()) for f in re.split(r"[;\n]", body) if "::" in f]


def test_header_parser_sees_the_whole_abi():
    protos = header_prototypes()
    assert len(protos) >= 60 and protos["esimp_create"] == ("int", ["esimp_ctx**", "int"])
    assert protos["esimp_last_error"] == ("char*", ["esimp_ctx*"])
    assert protos["esimp_multi_set_grid_structured"][1] == ["esimp_multi*", "int", "int", "int", "double", "double", "double", "int64_t*"]
    lib = C.CDLL(str(ROOT / "easysimp.jl_b200" / "libeasysimp_b200.so")) if (ROOT / "easysimp.jl_b200" / "libeasysimp_b200.so").exists() else None
    if lib is not None:
        for name in protos:
            assert hasattr(lib, name), f"{name} is declared but not exported"


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 35
    used = set()
    for name, ret, types, args, line in calls:
        where = f"EasySIMPB200.jl:{line} ccall(:{name})"
        assert name in protos, f"{where}: not declared in include/easysimp_b200.h"
        c_ret, c_params = protos[name]
        assert len(types) == len(c_params), f"{where}: {len(types)} argument types for {len(c_params)} C parameters"
        assert len(args) == len(c_params), f"{where}: {len(args)} arguments for {len(c_params)} C parameters"
        assert ret in C_TO_JULIA[c_ret], f"{where}: returns {ret}, the header says {c_ret}"
        for i, (jt, ct) in enumerate(zip(types, c_params)):
            assert jt in C_TO_JULIA[ct], f"{where}: argument {i + 1} is {jt}, the header says {ct}"
        used.add(name)
    # the binding covers the loop, the setup calls and the single-process multi-GPU handle
    for needed in ("esimp_create", "esimp_destroy", "esimp_set_params", "esimp_set_grid_structured", "esimp_set_mesh_unstructured",
                   "esimp_set_dirichlet", "esimp_set_load", "esimp_set_density", "esimp_matvec", "esimp_filter_build_structured",
                   "esimp_filter_build_csr", "esimp_init_design", "esimp_iterate", "esimp_final_analysis", "esimp_get_physical_density",
                   "esimp_get_displacement", "esimp_stress_fields", "esimp_stress_tensors", "esimp_set_body_force",
                   "esimp_multi_create", "esimp_multi_iterate", "esimp_multi_final_analysis", "esimp_multi_destroy"):
        assert needed in used, f"the binding never calls {needed}"


@pytest.mark.parametrize("c_name,j_name", [("esimp_params", "EsimpParams"), ("esimp_iter_stats", "EsimpIterStats")])
def test_structs_crossing_the_boundary_have_the_header_layout(c_name, j_name):
    cf, jf = header_struct(c_name), julia_struct(j_name)
    assert [n for n, _ in cf] == [n for n, _ in jf], f"{j_name}: field order differs from {c_name}"
    for (n, ct), (_, jt) in zip(cf, jf):
        assert jt in C_FIELD_TO_JULIA[ct], f"{j_name}.{n} is {jt}, the header says {ct}"
    # ... and the ctypes mirror agrees on the size, so all three statements of the struct are the same bytes
    import easysimp_b200.lib as lib

    size = sum({"double": 8, "int32_t": 4, "int64_t": 8}[t] for _, t in cf)
    assert C.sizeof({"esimp_params": lib.EsimpParams, "esimp_iter_stats": lib.EsimpIterStats}[c_name]) == size and size % 8 == 0


def test_binding_defines_the_reference_entry_points():
    """north_star: the Julia host keeps simp_optimize, OptimizationParameters, FilterCache / create_filter_cache and the
    assemble_stiffness_matrix_simp! entry points as a drop-in (reference: FiniteElementAnalysis.jl:130-157, :229, :293-340, :637;
    Optimization.jl:178)."""
    text = JULIA.read_text()
    for pattern in (r"function\s+setup_problem\w*\(", r"assemble_stiffness_matrix_simp!\(", r"apply_fixed_boundary!\(", r"apply_sliding_boundary!\(",
                    r"function\s+simp_optimize\w*\(", r"OptimizationParameters", r"create_filter_cache\(", r"tolerance_checkpoints",
                    r"export_interval", r"OptimizationResult\("):
        assert re.search(pattern, text), pattern
    # 1-based Julia indices are shifted exactly where dofs / cells cross the boundary
    assert re.search(r"prescribed[^\n]*\.- 1", text) and "cell[] + 1" in text


def test_julia_source_block_structure_balances():
    """No Julia parser here; the cheapest structural check that still catches a lost `end` or bracket: with strings, docstrings,
    comments and index expressions removed, block openers and `end`s pair up, and (), [], {} balance."""
    t = JULIA.read_text()
    t = re.sub(r'"""(.*?)"""', '""', t, flags=re.S)
    t = re.sub(r'"(\\.|[^"\\])*"', '""', t)
    t = re.sub(r"#[^\n]*", "", t)
    for a, b in ("()", "[]", "{}"):
        assert t.count(a) == t.count(b), f"unbalanced {a}{b}"
    out, depth = [], 0
    for ch in t:  # `end` inside [...] is an index, comprehension `for`s live there too
        depth += ch == "["
        if depth == 0:
            out.append(ch)
        depth -= ch == "]"
    t = "".join(out)
    openers = re.findall(r"(?<![\w.:])(function|if|for|while|struct|let|do|begin|try|module|quote|macro)(?![\w!])", t)
    ends = re.findall(r"(?<![\w.:])end(?![\w!])", t)
    assert len(openers) == len(ends) and len(openers) > 40
