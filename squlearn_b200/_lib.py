"""ctypes binding of libb200sq.so (the C ABI declared in include/b200sq.h).

There is deliberately no fallback: if the shared library is missing, ``load()`` raises, and every
compute entry point of the library itself fails when no CUDA device is present.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SQ_LIB", os.path.join(_HERE, "libb200sq.so"))

GATE_KINDS = [
    "id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx",
    "rx", "ry", "rz", "p", "u",
    "cx", "cy", "cz", "ch", "cs", "csx",
    "cp", "crx", "cry", "crz",
    "swap", "iswap", "ecr",
    "rxx", "ryy", "rzz", "rzx",
    "ccx", "cswap",
]
GATE_KIND = {name: i for i, name in enumerate(GATE_KINDS)}
# (number of qubits, takes an angle)
GATE_ARITY = {
    "id": (1, False), "h": (1, False), "x": (1, False), "y": (1, False), "z": (1, False),
    "s": (1, False), "sdg": (1, False), "t": (1, False), "tdg": (1, False), "sx": (1, False),
    "rx": (1, True), "ry": (1, True), "rz": (1, True), "p": (1, True), "u": (1, True),
    "cx": (2, False), "cy": (2, False), "cz": (2, False), "ch": (2, False), "cs": (2, False),
    "csx": (2, False), "cp": (2, True), "crx": (2, True), "cry": (2, True), "crz": (2, True),
    "swap": (2, False), "iswap": (2, False), "ecr": (2, False),
    "rxx": (2, True), "ryy": (2, True), "rzz": (2, True), "rzx": (2, True),
    "ccx": (3, False), "cswap": (3, False),
}

FN_CONST, FN_LINEAR, FN_ARCCOS, FN_ARCTAN, FN_TABLE = range(5)

GRAM_SYMMETRIC, GRAM_DIAG_ONE, GRAM_DIAG_ZERO, GRAM_3M, GRAM_CLASSIC, GRAM_INT8, GRAM_CRT = 1, 2, 4, 8, 16, 32, 64


class Gate(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("q", ctypes.c_int32 * 3),
        ("fn", ctypes.c_int32),
        ("feat", ctypes.c_int32),
        ("slot", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("c0", ctypes.c_double),
        ("c1", ctypes.c_double),
        ("c2", ctypes.c_double),
        ("c3", ctypes.c_double),
    ]


class GramJob(ctypes.Structure):
    """struct b200sq_gram_job (include/b200sq.h)."""
    _fields_ = [
        ("x0", ctypes.c_int64), ("nx", ctypes.c_int64),
        ("planes_y_dev", ctypes.c_void_p), ("expo_y_dev", ctypes.c_void_p),
        ("plane_rows_y", ctypes.c_int64), ("y0", ctypes.c_int64), ("ny", ctypes.c_int64),
        ("K_dev", ctypes.c_void_p), ("ldk", ctypes.c_int64),
        ("flags", ctypes.c_uint32), ("ready_value", ctypes.c_uint32),
        ("ready_flag_dev", ctypes.c_void_p),
        ("align_yx_dev", ctypes.c_void_p), ("align_yy_dev", ctypes.c_void_p),
        ("align_out_dev", ctypes.c_void_p), ("align_weight", ctypes.c_double),
        ("K_mirror_dev", ctypes.c_void_p), ("ldk_mirror", ctypes.c_int64),
    ]


EXPORTS = [
    "b200sq_last_error", "b200sq_abi_version", "b200sq_device_count", "b200sq_program_create",
    "b200sq_program_destroy", "b200sq_program_set_coeffs", "b200sq_program_num_passes",
    "b200sq_program_dump", "b200sq_program_jit_compile", "b200sq_simulate", "b200sq_simulate_rowmax",
    "b200sq_gram_rowmax", "b200sq_gram_slice_rowmax", "b200sq_simulate_expvals2", "b200sq_gram_crt_moduli", "b200sq_gram_crt_bits",
    "b200sq_gram_slice_crt", "b200sq_gram_slice_crt_digits", "b200sq_gram_digits_to_crt", "b200sq_simulate_expvals", "b200sq_expvals",
    "b200sq_gram", "b200sq_gram_planes_bytes", "b200sq_gram_slice", "b200sq_gram_from_planes", "b200sq_gram_jobs", "b200sq_wait_flags", "b200sq_gram_ready_timeouts", "b200sq_ipc_export", "b200sq_ipc_open", "b200sq_ipc_close",
    "b200sq_copy_async", "b200sq_copy_planes_async", "b200sq_rbf",
    "b200sq_launch_count",
]

_lib = None


def load():
    """Load libb200sq.so and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA engine with `python -m squlearn_b200.build` "
            "(squlearn_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, i64, dbl = c.c_void_p, c.c_int32, c.c_int64, c.c_double
    lib.b200sq_last_error.restype = c.c_char_p
    lib.b200sq_last_error.argtypes = []
    lib.b200sq_abi_version.restype = c.c_int
    lib.b200sq_device_count.restype = c.c_int
    lib.b200sq_launch_count.restype = i64
    lib.b200sq_program_create.restype = c.c_int
    lib.b200sq_program_create.argtypes = [c.c_int, c.c_int, c.POINTER(Gate), c.c_int, c.POINTER(vp)]
    lib.b200sq_program_destroy.restype = None
    lib.b200sq_program_destroy.argtypes = [vp]
    lib.b200sq_program_set_coeffs.restype = c.c_int
    lib.b200sq_program_set_coeffs.argtypes = [vp, c.POINTER(dbl), c.POINTER(dbl)]
    lib.b200sq_program_num_passes.restype = c.c_int
    lib.b200sq_program_num_passes.argtypes = [vp]
    lib.b200sq_program_dump.restype = c.c_int
    lib.b200sq_program_dump.argtypes = [vp, c.POINTER(i32), c.c_int, c.POINTER(c.c_int)]
    lib.b200sq_program_jit_compile.restype = c.c_long
    lib.b200sq_program_jit_compile.argtypes = [vp, c.c_int, c.c_char_p, c.c_int]
    lib.b200sq_simulate.restype = c.c_int
    lib.b200sq_simulate.argtypes = [vp, i64, c.c_int, vp, vp, c.c_int, c.POINTER(i32),
                                    c.POINTER(dbl), vp, vp]
    lib.b200sq_simulate_rowmax.restype = c.c_int
    lib.b200sq_simulate_rowmax.argtypes = [vp, i64, c.c_int, vp, vp, c.c_int, c.POINTER(i32),
                                           c.POINTER(dbl), vp, vp, vp]
    lib.b200sq_gram_rowmax.restype = c.c_int
    lib.b200sq_gram_rowmax.argtypes = [vp, i64, vp, i64, i64, vp, i64, c.c_uint32, vp, vp, vp]
    lib.b200sq_gram_slice_rowmax.restype = c.c_int
    lib.b200sq_gram_slice_rowmax.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp, vp]
    lib.b200sq_simulate_expvals.restype = c.c_int
    lib.b200sq_simulate_expvals.argtypes = [vp, i64, c.c_int, vp, vp, c.c_int, c.POINTER(i32),
                                            c.POINTER(dbl), c.c_int, c.POINTER(c.c_uint32),
                                            c.POINTER(c.c_uint32), vp, vp, c.c_size_t, vp]
    lib.b200sq_simulate_expvals2.restype = c.c_int
    lib.b200sq_simulate_expvals2.argtypes = [vp, i64, c.c_int, vp, vp, c.c_int, c.POINTER(i32), c.POINTER(dbl),
                                             c.POINTER(i32), c.POINTER(dbl), c.c_int, c.POINTER(c.c_uint32),
                                             c.POINTER(c.c_uint32), vp, vp, c.c_size_t, vp]
    lib.b200sq_expvals.restype = c.c_int
    lib.b200sq_expvals.argtypes = [vp, i64, c.c_int, c.c_int, c.POINTER(c.c_uint32),
                                   c.POINTER(c.c_uint32), vp, vp]
    lib.b200sq_gram.restype = c.c_int
    lib.b200sq_gram.argtypes = [vp, i64, vp, i64, i64, vp, i64, c.c_uint32, vp]
    lib.b200sq_gram_planes_bytes.restype = c.c_size_t
    lib.b200sq_gram_planes_bytes.argtypes = [i64, i64]
    lib.b200sq_gram_slice.restype = c.c_int
    lib.b200sq_gram_slice.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp]
    lib.b200sq_gram_from_planes.restype = c.c_int
    lib.b200sq_gram_from_planes.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64, i64, i64, i64, vp, i64, c.c_uint32, vp]
    lib.b200sq_gram_crt_moduli.restype = c.c_int
    lib.b200sq_gram_crt_moduli.argtypes = [i64]
    lib.b200sq_gram_crt_bits.restype = c.c_int
    lib.b200sq_gram_crt_bits.argtypes = [i64]
    lib.b200sq_gram_slice_crt.restype = c.c_int
    lib.b200sq_gram_slice_crt.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp, vp]
    lib.b200sq_gram_slice_crt_digits.restype = c.c_int
    lib.b200sq_gram_slice_crt_digits.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp, vp, i64, i64, vp]
    lib.b200sq_gram_digits_to_crt.restype = c.c_int
    lib.b200sq_gram_digits_to_crt.argtypes = [vp, i64, i64, i64, i64, vp, i64, i64, vp]
    lib.b200sq_gram_jobs.restype = c.c_int
    lib.b200sq_gram_jobs.argtypes = [vp, vp, i64, i64, c.c_int, c.POINTER(GramJob), vp]
    lib.b200sq_wait_flags.restype = c.c_int
    lib.b200sq_wait_flags.argtypes = [vp, c.POINTER(i32), c.c_int, c.c_uint32, vp]
    lib.b200sq_gram_ready_timeouts.restype = i64
    lib.b200sq_gram_ready_timeouts.argtypes = []
    lib.b200sq_ipc_export.restype = c.c_int
    lib.b200sq_ipc_export.argtypes = [vp, c.c_char_p, c.POINTER(i64)]
    lib.b200sq_ipc_open.restype = c.c_int
    lib.b200sq_ipc_open.argtypes = [c.c_char_p, c.POINTER(vp)]
    lib.b200sq_ipc_close.restype = c.c_int
    lib.b200sq_ipc_close.argtypes = [vp]
    lib.b200sq_copy_async.restype = c.c_int
    lib.b200sq_copy_async.argtypes = [vp, vp, c.c_size_t, vp]
    lib.b200sq_copy_planes_async.restype = c.c_int
    lib.b200sq_copy_planes_async.argtypes = [vp, c.c_size_t, vp, c.c_size_t, c.c_size_t, c.c_int, vp]
    lib.b200sq_rbf.restype = c.c_int
    lib.b200sq_rbf.argtypes = [vp, i64, vp, i64, c.c_int, dbl, vp, i64, vp]
    if lib.b200sq_abi_version() != 3:
        raise RuntimeError("libb200sq.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(code):
    """Map a non-zero status to the exception classes the reference treats as critical
    (never retried): RuntimeError / ValueError (src/squlearn/util/executor.py:1054-1099)."""
    if code == 0:
        return
    msg = load().b200sq_last_error().decode("utf-8", "replace")
    if code == 2:
        raise ValueError(f"b200sq: {msg}")
    raise RuntimeError(f"b200sq error {code}: {msg}")
