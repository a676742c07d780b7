"""Shared test helpers: random circuits expressed both as a product Program and as an oracle gate list."""

import numpy as np

import oracle
from oracle.statevector import Gate
from squlearn_b200.program import Angle, GateOp, Program

FIXED_1Q = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "id"]
ROT_1Q = ["rx", "ry", "rz", "p"]
FIXED_2Q = ["cx", "cy", "cz", "ch", "cs", "csx", "swap", "iswap", "ecr"]
ROT_2Q = ["cp", "crx", "cry", "crz", "rxx", "ryy", "rzz", "rzx"]
FIXED_3Q = ["ccx", "cswap"]


def _oracle_angle(a: Angle):
    fn = {None: None, "linear": (lambda v: v), "arccos": np.arccos, "arctan": np.arctan}[a.fn]

    def f(x, p, a=a, fn=fn):
        val = a.scale
        if a.param is not None:
            val = val * p[a.param]
        if a.feature is not None:
            val = val * fn(x[:, a.feature])
        return a.offset + val

    return f


def random_circuit(n, n_gates, num_params, num_features, rng, kinds=None):
    """Random circuit over the whole gate vocabulary. Returns (Program, oracle gate list)."""
    ops, ogates = [], []
    pool = kinds or (FIXED_1Q + ROT_1Q + ["u"] + (FIXED_2Q + ROT_2Q if n >= 2 else [])
                     + (FIXED_3Q if n >= 3 else []))
    for _ in range(n_gates):
        name = pool[rng.integers(len(pool))]
        nq = 1 if name in FIXED_1Q + ROT_1Q + ["u"] else (3 if name in FIXED_3Q else 2)
        qs = tuple(int(q) for q in rng.choice(n, size=nq, replace=False))
        angle, extra = None, (0.0, 0.0)
        if name in ROT_1Q + ROT_2Q + ["u"]:
            mode = rng.integers(4)
            scale = float(rng.uniform(0.5, 1.5))
            offset = float(rng.uniform(-0.3, 0.3)) if rng.integers(2) else 0.0
            if mode == 0:
                angle = Angle(param=int(rng.integers(num_params)), scale=scale, offset=offset)
            elif mode == 1:
                angle = Angle(feature=int(rng.integers(num_features)), fn="linear", scale=scale,
                              offset=offset)
            elif mode == 2:
                angle = Angle(param=int(rng.integers(num_params)),
                              feature=int(rng.integers(num_features)),
                              fn=["arccos", "arctan", "linear"][rng.integers(3)], scale=scale)
            else:
                angle = Angle(scale=scale, offset=offset)  # constant
            if name == "u":
                extra = (float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        ops.append(GateOp(name, qs, angle, extra))
        if name == "u":
            ogates.append(Gate("u", qs, (_oracle_angle(angle), extra[0], extra[1])))
        elif angle is not None:
            ogates.append(Gate(name, qs, (_oracle_angle(angle),)))
        else:
            ogates.append(Gate(name, qs, ()))
    return Program(n, ops, num_params, num_features), ogates


def program_to_oracle(program: Program):
    """Oracle gate list of a product Program (used where the oracle has no builder of its own)."""
    out = []
    for g in program.gates:
        if g.name == "u":
            out.append(Gate("u", g.qubits, (_oracle_angle(g.angle), g.extra[0], g.extra[1])))
        elif g.angle is not None:
            out.append(Gate(g.name, g.qubits, (_oracle_angle(g.angle),)))
        else:
            out.append(Gate(g.name, g.qubits, ()))
    return out


# ------------------------------------------------------------------------------------------------
# NumPy restatement of the split-integer Gram (csrc/gram_i8.cu): digit planes and their recombination.
# Test infrastructure: lets the CPU suite check the plane layout / exchange logic and the error of the
# scheme without a GPU.
I8_SLICES = 6


def i8_slice(psi):
    """(planes [12, N, 2 D] int8, expo [N] int32) of complex128 states, like k_slice_i8."""
    psi = np.asarray(psi, dtype=np.complex128)
    n_rows, dim = psi.shape
    nat = np.stack([psi.real, psi.imag], axis=2).reshape(n_rows, 2 * dim)
    rot = np.stack([-psi.imag, psi.real], axis=2).reshape(n_rows, 2 * dim)
    m = np.abs(nat).max(axis=1)
    expo = np.where(m > 0, np.frexp(m)[1], 0).astype(np.int32)
    planes = np.zeros((2 * I8_SLICES, n_rows, 2 * dim), dtype=np.int8)
    for base, v in ((0, nat), (I8_SLICES, rot)):
        q = np.rint(np.ldexp(v, (6 + 8 * (I8_SLICES - 1) - expo)[:, None])).astype(np.int64)
        for s_ in range(I8_SLICES - 1, 0, -1):
            d = ((q + 128) & 255) - 128
            q = (q - d) >> 8
            planes[base + s_] = d.astype(np.int8)
        planes[base] = q.astype(np.int8)
    return planes, expo


def i8_gram(planes_x, expo_x, planes_y, expo_y):
    """|<x|y>|^2 from digit planes: classes c <= 5 plus the pair (3, 3), like k_gram_i8 + k_gram_i8_final.
    Re = nat(x) . nat(y), Im = rot(x) . nat(y): the column operand only needs its 6 nat planes."""
    S = I8_SLICES
    px = planes_x.astype(np.int64)
    py = planes_y.astype(np.int64)
    pairs = [(6, [(3, 3)])] + [(c, [(i, c - i) for i in range(c + 1)]) for c in range(S - 1, -1, -1)]
    re = np.zeros((px.shape[1], py.shape[1]))
    im = np.zeros_like(re)
    for c, lst in pairs:
        ar = sum(px[i] @ py[j].T for i, j in lst)
        ai = sum(px[S + i] @ py[j].T for i, j in lst)
        re += ar * 2.0 ** (-8 * c)
        im += ai * 2.0 ** (-8 * c)
    scale = np.ldexp(1.0, 2 * (expo_x[:, None].astype(np.int64) + expo_y[None, :] - 12))
    return (re * re + im * im) * scale


# ------------------------------------------------------------------------------------------------
# NumPy restatement of the CRT form of the split-integer Gram (csrc/gram_i8.cu, B200SQ_GRAM_CRT): residue planes of
# q = round(v * 2^(B - e)) modulo pairwise coprime p_m <= 256 (256 and odd numbers), e the smallest integer with
# 2^(B - e) |row|_2 <= R_eff ~ sqrt(P / 2) (Cauchy-Schwarz keeps every inner product below P / 2), one exact integer
# product per modulus, Chinese remainder reconstruction with the kernel's double-double accumulation of k_m / p_m.
CRT_MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193]
CRT_MAX_BITS = 46
CRT_MAX_LOG2R = 46.9
CRT_TOL = 1e-10


def crt_plan(dim):
    """(number of moduli, B, e_bias) like crt_plan() of gram_i8.cu."""
    from decimal import Decimal, getcontext

    getcontext().prec = 60
    kp = 2 * dim
    need = int(np.ceil(np.log2(2.0 * np.sqrt(kp) / CRT_TOL)))
    big = 1
    for s_ in range(1, len(CRT_MODULI) + 1):
        big *= CRT_MODULI[s_ - 1]
        r_eff = (Decimal(big) / 2).sqrt() - Decimal("0.5") * Decimal(kp).sqrt() - 1
        if r_eff < 2:
            continue
        log2r = min(float(r_eff.ln() / Decimal(2).ln()), CRT_MAX_LOG2R)
        bits = min(int(np.floor(log2r)), CRT_MAX_BITS)
        if bits >= need or (s_ == len(CRT_MODULI) and bits >= 30):
            return s_, bits, bits - log2r
    raise AssertionError("no CRT plan")


def crt_num_moduli(dim):
    return crt_plan(dim)[0]


def _symres(q, p):
    r = np.mod(q, p)
    return np.where(r > (p - 1) // 2, r - p, r)   # (p = 256: -128 .. 127)


def crt_expo(psi, dim):
    _, bits, e_bias = crt_plan(dim)
    n2 = np.sum(np.abs(np.asarray(psi, dtype=np.complex128)) ** 2, axis=1)
    return np.where(n2 > 0, np.ceil(0.5 * np.log2(np.where(n2 > 0, n2, 1.0)) + e_bias + 1e-9), 0).astype(np.int32)


def crt_slice(psi):
    """(planes [2 s, N, 2 D] int8, expo [N] int32) like k_slice_crt."""
    psi = np.asarray(psi, dtype=np.complex128)
    n_rows, dim = psi.shape
    s_, bits, _ = crt_plan(dim)
    nat = np.stack([psi.real, psi.imag], axis=2).reshape(n_rows, 2 * dim)
    expo = crt_expo(psi, dim)
    q = np.rint(np.ldexp(nat, (bits - expo)[:, None])).astype(np.int64)
    rot = np.empty_like(q)
    rot[:, 0::2], rot[:, 1::2] = -q[:, 1::2], q[:, 0::2]
    planes = np.zeros((2 * s_, n_rows, 2 * dim), dtype=np.int8)
    for mi in range(s_):
        planes[mi] = _symres(q, CRT_MODULI[mi]).astype(np.int8)
        planes[s_ + mi] = _symres(rot, CRT_MODULI[mi]).astype(np.int8)
    return planes, expo


def crt_digits(psi):
    """(digits [6, N, 2 D] int8, expo [N] int32): the nat digit planes k_slice_crt writes for the multi-GPU exchange,
    balanced base-256 digits of the CRT integers q = round(v * 2^(B - e)) themselves."""
    psi = np.asarray(psi, dtype=np.complex128)
    n_rows, dim = psi.shape
    nat = np.stack([psi.real, psi.imag], axis=2).reshape(n_rows, 2 * dim)
    expo = crt_expo(psi, dim)
    q = np.rint(np.ldexp(nat, (crt_plan(dim)[1] - expo)[:, None])).astype(np.int64)
    planes = np.zeros((I8_SLICES, n_rows, 2 * dim), dtype=np.int8)
    for s_ in range(I8_SLICES - 1, 0, -1):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        planes[s_] = d.astype(np.int8)
    planes[0] = q.astype(np.int8)
    return planes, expo


def crt_gram(planes_x, expo_x, planes_y, expo_y):
    """|<x|y>|^2 from residue planes like k_gram_i8 (CRT classes) + k_gram_crt_final."""
    s_ = planes_x.shape[0] // 2
    mods = CRT_MODULI[:s_]
    big = 1
    for p in mods:
        big *= p
    from fractions import Fraction

    bits = crt_plan(planes_x.shape[2] // 2)[1]
    scale = float(Fraction(big, 2 ** (2 * bits)))
    parts = []
    for half in (0, s_):
        sh = np.zeros((planes_x.shape[1], planes_y.shape[1]))
        sl = np.zeros_like(sh)
        for mi, p in enumerate(mods):
            a = planes_x[half + mi].astype(np.int64)
            b = planes_y[mi].astype(np.int64)
            r = _symres(a @ b.T, p)
            c = pow(big // p, -1, p)
            k = np.mod(r * c, p).astype(np.float64)
            inv = 1.0 / p
            hi = k * inv
            # fma(-hi, p, k) on the device; hi * p has at most 61 significant bits, so x87 long double is exact
            lo = (k.astype(np.longdouble) - hi.astype(np.longdouble) * np.longdouble(p)).astype(np.float64) * inv
            t = sh + hi
            bb = t - sh
            sl += ((sh - (t - bb)) + (hi - bb)) + lo
            sh = t
        f = (sh - np.rint(sh)) + sl
        f = np.where(f > 0.5, f - 1.0, f)
        f = np.where(f < -0.5, f + 1.0, f)
        parts.append(f * scale)
    re, im = parts
    return (re * re + im * im) * np.ldexp(1.0, 2 * (expo_x[:, None].astype(np.int64) + expo_y[None, :]))
