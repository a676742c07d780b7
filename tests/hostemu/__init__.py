"""TEST INFRASTRUCTURE: ctypes wrapper of the host emulator (see hostemu.cpp)."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libb200sq_hostemu.so")
SRC = os.path.join(HERE, "hostemu.cpp")
DEPS = [SRC, os.path.join(ROOT, "squlearn_b200", "csrc", "plan.hpp"),
        os.path.join(ROOT, "squlearn_b200", "csrc", "tile_exec.cuh"),
        os.path.join(ROOT, "squlearn_b200", "csrc", "plan_dev.h"),
        os.path.join(ROOT, "include", "b200sq.h")]


def build(force=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in DEPS):
        return SO
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "include"), "-o", SO, SRC]
    subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostemu_last_error.restype = ctypes.c_char_p
        _lib.hostemu_simulate.restype = ctypes.c_int
    return _lib


def simulate(program, params, x, tile_bits=0, variants=None, team_size=32):
    """Run the product's plan + micro-op interpreter on the CPU. Returns (states, n_passes)."""
    from squlearn_b200 import _lib as blib  # only for the Gate struct definition

    x = np.ascontiguousarray(x, dtype=np.float64)
    n_samples, d = x.shape
    gates = program.c_gates(params)
    table = program.angle_table(x, params)
    vg2 = vs2 = None
    if variants is None:
        nv, vg, vs = 1, None, None
    else:
        vg_arr = np.ascontiguousarray(variants[0], dtype=np.int32)
        vs_arr = np.ascontiguousarray(variants[1], dtype=np.float64)
        nv = len(vg_arr)
        vg = vg_arr.ctypes.data_as(ctypes.c_void_p)
        vs = vs_arr.ctypes.data_as(ctypes.c_void_p)
        if len(variants) > 2:   # second shift per variant (second-order derivatives)
            vg2_arr = np.ascontiguousarray(variants[2], dtype=np.int32)
            vs2_arr = np.ascontiguousarray(variants[3], dtype=np.float64)
            vg2 = vg2_arr.ctypes.data_as(ctypes.c_void_p)
            vs2 = vs2_arr.ctypes.data_as(ctypes.c_void_p)
    dim = 1 << program.num_qubits
    psi = np.zeros((nv, n_samples, dim), dtype=np.complex128)
    rc = lib().hostemu_simulate(
        ctypes.c_int(program.num_qubits), ctypes.c_int(len(program)), gates,
        ctypes.c_int(tile_bits), ctypes.c_int64(n_samples), ctypes.c_int(d),
        x.ctypes.data_as(ctypes.c_void_p),
        table.ctypes.data_as(ctypes.c_void_p) if table is not None else None,
        ctypes.c_int(nv), vg, vs, psi.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(team_size), vg2, vs2)
    if rc < 0:
        raise RuntimeError(lib().hostemu_last_error().decode())
    return (psi if variants is not None else psi[0]), rc


def expvals(psi, n_qubits, xmasks, zmasks):
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    n_states = psi.size >> n_qubits
    xm = np.ascontiguousarray(xmasks, dtype=np.uint32)
    zm = np.ascontiguousarray(zmasks, dtype=np.uint32)
    out = np.zeros((n_states, len(xm)))
    lib().hostemu_expvals(psi.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n_states),
                          ctypes.c_int(n_qubits), ctypes.c_int(len(xm)),
                          xm.ctypes.data_as(ctypes.c_void_p), zm.ctypes.data_as(ctypes.c_void_p),
                          out.ctypes.data_as(ctypes.c_void_p))
    return out
