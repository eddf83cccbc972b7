"""include/simt_b200.h against the ctypes table of simt_b200/_lib.py: every prototype's parameter COUNT and KINDS
(pointer / int / long long / size_t / float) must match the argtypes the Python side binds -- a drifted signature would
only show on a GPU (as a crash or silently shifted arguments).  Also checks the call sites of HeadRunner against the
table through ctypes' own argument conversion.  No GPU."""
import ctypes
import os
import re

from simt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prototypes():
    text = open(os.path.join(ROOT, "include", "simt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(simt_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def kind_of_c(decl):
    d = decl.replace("const", " ").strip()
    if "*" in d:
        return "ptr"
    base = re.sub(r"\b\w+$", "", d).strip() if len(d.split()) > 1 else d      # drop the parameter name
    base = " ".join(base.split())
    return {"int": "int", "long long": "ll", "size_t": "size", "float": "float", "double": "double"}.get(base, base)


def kind_of_ctypes(t):
    if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    return {ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_size_t: "size", ctypes.c_float: "float",
            ctypes.c_double: "double"}[t]


def test_every_prototype_matches_its_ctypes_signature():
    protos = prototypes()
    assert len(protos) >= 30
    missing = sorted(set(protos) - set(_lib.SIGNATURES))
    extra = sorted(set(_lib.SIGNATURES) - set(protos))
    assert not missing, f"declared in the header but not bound: {missing}"
    assert not extra, f"bound but not declared in the header: {extra}"
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, _lib.py binds {len(args)}"
        for i, (p, a) in enumerate(zip(params, args)):
            assert kind_of_c(p) == kind_of_ctypes(a), f"{name} parameter {i} ({p!r}): header {kind_of_c(p)}, ctypes {kind_of_ctypes(a)}"
        rk = "void" if ret.replace("const", "").strip() == "void" else ("ptr" if "*" in ret else kind_of_c(ret + " x"))
        assert rk == ("void" if res is None else kind_of_ctypes(res)), f"{name}: return type {ret!r} vs {res}"


def test_every_python_call_site_passes_the_bound_number_of_arguments():
    """ctypes accepts surplus arguments for cdecl functions and only complains about missing ones at call time (on a
    GPU box); count the positional arguments of every `<something>.simt_xxx(...)` call in the Python sources instead."""
    import ast
    import glob
    files = (glob.glob(os.path.join(ROOT, "simt_b200", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py")) +
             glob.glob(os.path.join(ROOT, "scripts", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    seen, bad = 0, []
    for path in files:
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in _lib.SIGNATURES:
                if any(isinstance(a, ast.Starred) for a in node.args) or node.keywords:
                    continue
                seen += 1
                want = len(_lib.SIGNATURES[node.func.attr][1])
                if len(node.args) != want:
                    bad.append(f"{os.path.relpath(path, ROOT)}:{node.lineno}: {node.func.attr} called with {len(node.args)} arguments, bound with {want}")
    assert seen >= 40, seen
    assert not bad, "\n".join(bad)
