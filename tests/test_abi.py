"""CPU tests of the C-ABI library: it loads, exports every symbol include/b200sq.h declares, plans
without a GPU, and FAILS LOUDLY (no fallback) when asked to compute without one."""

import ctypes
import os
import re

import numpy as np
import pytest

from squlearn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200sq.h")).read()
    return sorted(set(re.findall(r"B200SQ_API[^;(]*?\b(b200sq_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) == len(_lib.EXPORTS) >= 23
    assert sorted(names) == sorted(_lib.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert lib.b200sq_abi_version() == 3


def test_gate_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Gate) == 64
    assert _lib.Gate.c0.offset == 32 and _lib.Gate.c3.offset == 56
    assert len(_lib.GATE_KINDS) == 34  # B200SQ_NUM_GATE_KINDS
    text = open(os.path.join(ROOT, "include", "b200sq.h")).read()
    enum = re.search(r"enum b200sq_gate_kind \{(.*?)B200SQ_NUM_GATE_KINDS", text, re.S).group(1)
    header_names = [m.lower() for m in re.findall(r"B200SQ_(\w+)", enum)]
    assert header_names == _lib.GATE_KINDS


def test_gram_job_struct_layout_matches_header():
    """struct b200sq_gram_job: field order and size as declared in include/b200sq.h."""
    text = open(os.path.join(ROOT, "include", "b200sq.h")).read()
    body = re.search(r"typedef struct b200sq_gram_job \{(.*?)\} b200sq_gram_job;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip().lstrip("*") for n in decl.split(",")]
            names[-len(decl.split(",")):] = [re.findall(r"(\w+)$", n.strip())[0] for n in decl.split(",")]
    assert names == [f[0] for f in _lib.GramJob._fields_]
    assert ctypes.sizeof(_lib.GramJob) == 136
    assert _lib.GramJob.ready_flag_dev.offset == 80 and _lib.GramJob.align_weight.offset == 112


def test_plan_without_gpu_and_error_reporting():
    lib = _lib.load()
    gates = (_lib.Gate * 2)()
    gates[0].kind, gates[0].q[0], gates[0].q[1], gates[0].q[2] = _lib.GATE_KIND["h"], 0, -1, -1
    gates[1].kind, gates[1].q[0], gates[1].q[1], gates[1].q[2] = _lib.GATE_KIND["cx"], 0, 1, -1
    h = ctypes.c_void_p()
    _lib.check(lib.b200sq_program_create(2, 2, gates, 0, ctypes.byref(h)))
    assert lib.b200sq_program_num_passes(h) == 1
    n = ctypes.c_int(0)
    _lib.check(lib.b200sq_program_dump(h, None, 0, ctypes.byref(n)))
    assert n.value > 5
    lib.b200sq_program_destroy(h)
    # bad qubit -> ValueError with a message
    gates[1].q[1] = 7
    with pytest.raises(ValueError, match="qubit"):
        _lib.check(lib.b200sq_program_create(2, 2, gates, 0, ctypes.byref(h)))


@pytest.mark.skipif(_lib.load().b200sq_device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device every product entry point raises instead of computing on the CPU."""
    import squlearn_b200 as sq

    with pytest.raises(RuntimeError, match="CUDA"):
        sq.Executor("b200")
    lib = _lib.load()
    buf = np.zeros(16, dtype=np.complex128)
    out = np.zeros(4)
    rc = lib.b200sq_gram(buf.ctypes.data, 2, buf.ctypes.data, 2, 2, out.ctypes.data, 2, 0, None)
    assert rc != 0
    assert b"no CUDA device" in lib.b200sq_last_error() or b"CPU path" in lib.b200sq_last_error()
    with pytest.raises(ValueError, match="Unknown backend string"):
        sq.Executor("pennylane")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "squlearn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert not re.search(r"^\s*(from|import)\s+tests\b", src, re.M), f


def test_crt_plan_of_the_library_equals_the_restatement():
    """The host-side CRT plan (number of moduli and bits per state dimension, gram_i8.cu: crt_plan) equals the NumPy
    restatement the CPU error tests use (tests/helpers.py), for every dimension the INT8 Gram supports; the worst-case
    bound 2 sqrt(K') 2^-B of a unit-norm entry stays below 1e-10 wherever the moduli suffice."""
    from tests.helpers import crt_plan

    lib = _lib.load()
    for n in range(6, 23):
        s_, bits, _ = crt_plan(1 << n)
        assert lib.b200sq_gram_crt_moduli(1 << n) == s_, n
        assert lib.b200sq_gram_crt_bits(1 << n) == bits, n
        assert 2.0 * np.sqrt(2.0 ** (n + 1)) * 2.0 ** -bits <= 1e-10, n
    assert lib.b200sq_gram_crt_moduli(32) == 0          # below the smallest supported dimension
    assert (lib.b200sq_gram_crt_moduli(1 << 16), lib.b200sq_gram_crt_bits(1 << 16)) == (11, 43)   # BASELINE configs[2]


def test_row_stats_struct_layout():
    text = open(os.path.join(ROOT, "include", "b200sq.h")).read()
    body = re.search(r"typedef struct b200sq_row_stats \{(.*?)\} b200sq_row_stats;", text, re.S).group(1)
    assert [w for w in re.findall(r"(\w+);", body)] == ["norm2", "max_hi", "reserved"]
