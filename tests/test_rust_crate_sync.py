"""The pyo3 host crate shipped as source (rust/): not compilable in the build image (no cargo / rustc), so this test
keeps its FFI declarations in step with include/ssqcuda.h -- same entry points, same number of arguments, pointer
arguments where the header has pointers -- and checks that every `ffi::ssq_*` call in the #[pyfunction] bodies names a
declared entry point and passes the declared number of arguments, and that the module registers the reference's
callables (rust/src/lib.rs:25-32 of the reference)."""
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def _header():
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ssqcuda.h")).read(), flags=re.S)
    d = {}
    for m in re.finditer(r"\b(ssq_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = " ".join(m.group(2).split())
        al = [] if args in ("", "void") else _split_args(args)
        d[m.group(1)] = ["*" in a for a in al]
    return d


def _ffi():
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    blk = src[src.index('extern "C" {'):]
    blk = blk[:blk.index("\n}\n")]
    d = {}
    for m in re.finditer(r"pub fn (ssq_[a-z0-9_]+)\((.*?)\)(?:\s*->\s*[^;]+)?;", blk, flags=re.S):
        al = _split_args(m.group(2)) if m.group(2).strip() else []
        d[m.group(1)] = ["*" in a.split(":", 1)[1] for a in al]
    return d


def test_ffi_declarations_match_the_header():
    h, f = _header(), _ffi()
    assert set(h) == set(f), (sorted(set(h) - set(f)), sorted(set(f) - set(h)))
    for name in h:
        assert h[name] == f[name], (name, h[name], f[name])


def test_pyfunction_bodies_call_declared_entry_points_with_the_declared_arity():
    f = _ffi()
    n_calls = 0
    for fn in ("lib.rs", "spectral.rs", "wavelets.rs"):
        src = open(os.path.join(ROOT, "rust", "src", fn)).read()
        for m in re.finditer(r"ffi::(ssq_[a-z0-9_]+)\(", src):
            name = m.group(1)
            assert name in f, (fn, name)
            depth, i = 1, m.end()
            while depth:
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            args = _split_args(src[m.end():i - 1])
            assert len(args) == len(f[name]), (fn, name, len(args), len(f[name]))
            n_calls += 1
    assert n_calls >= 15


def test_module_registers_the_reference_callables():
    src = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    reg = set(re.findall(r"wrap_pyfunction!\((?:\w+::)?(\w+), py\)", src))
    for name in ("hello_from_bin", "stft", "ssq_stft", "cwt", "cwt_simd", "ssq_cwt"):  # reference lib.rs:25-32
        assert name in reg, name
    for name in ("istft", "issq_stft", "icwt", "morlet", "morlet_freq", "morlet_time", "gmw", "gmw_freq", "gmw_time",
                 "gmw_center_frequency"):  # north star + src/ssqueeze/_rs.pyi:61-132
        assert name in reg, name
    # and the Python mirror the tests drive exposes the same names
    from ssqueeze_rs_b200 import _rs
    for name in reg:
        assert hasattr(_rs, name), name
