"""The bindings a maintainer would add must match the header: every `@fb2` / `ccall` in the Julia shim and every ctypes
prototype of the Python mirror names a declared function and passes the declared number of arguments (CPU only)."""
import os
import re

import ferrite_b200 as fb
from ferrite_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_arity():
    src = open(os.path.join(ROOT, "include", "ferrite_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(fb2_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def split_top(s):
    """split a Julia tuple body at top-level commas"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return [p for p in parts if p.strip()]


def test_julia_shim_matches_header():
    ar = header_arity()
    src = open(os.path.join(ROOT, "ferrite.jl_b200", "julia", "FerriteB200.jl")).read()
    seen = 0
    for m in re.finditer(r"@fb2\s+(fb2_[a-z0-9_]+)\s+\(", src):
        name = m.group(1)
        # balanced tuple of argument types
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        ntypes = len(split_top(src[i:j - 1]))
        assert name in ar, name
        assert ntypes == ar[name], (name, ntypes, ar[name])
        seen += 1
    for m in re.finditer(r"ccall\(\(:(fb2_[a-z0-9_]+), LIB\), \w+, \(([^)]*)\)", src):
        name = m.group(1)
        assert name in ar, name
        assert len(split_top(m.group(2))) == ar[name], name
        seen += 1
    assert seen >= 15


def test_python_prototypes_match_header():
    ar = header_arity()
    protos = getattr(L, "_PROTOS", None)
    assert protos is not None
    for name, argtypes in protos.items():
        assert name in ar, name
        assert len(argtypes) == ar[name], (name, len(argtypes), ar[name])
    missing = [n for n in ar if n not in protos and n not in ("fb2_last_error", "fb2_version", "fb2_last_kernel")]
    assert missing == [], missing
