"""Gate lists of the encoding circuits on the hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Plain-function restatements (no Qiskit) of
  * ChebyshevPQC.get_circuit              src/squlearn/encoding_circuit/circuit_library/chebyshev_pqc.py:286-347
    parameter count :87-95, bounds :97-130, initial parameters :132-157
  * HubregtsenEncodingCircuit.get_circuit .../hubregtsen_encoding_circuit.py:125-190, bounds :68-88
  * ChebyshevRx.get_circuit               .../chebyshev_rx.py:203-269, bounds :64-77, initial parameters :79-103
  * EncodingCircuitBase.generate_initial_parameters   src/squlearn/encoding_circuit/encoding_circuit_base.py:74-92
  * LayeredEncodingCircuit Ry("p") / Rx("x") layers (running parameter index, features modulo d)
    .../layered_encoding_circuit.py:20-76,238-257 — as used by the reference's closed-form kernel tests
  * HighDimEncodingCircuit (columns layout, cx entangling, "saw" cycling)
    .../highdim_encoding_circuit.py:150-290
Each function returns a list of ``oracle.statevector.Gate``.
"""

import numpy as np

from .statevector import Gate


def _nonlin(name):
    if name == "arccos":
        return np.arccos
    if name == "arctan":
        return np.arctan
    raise ValueError(f"Unknown value for nonlinearity: {name}")


def _p(k):
    return lambda x, p, k=k: p[k]


def _x(f):
    return lambda x, p, f=f: x[:, f]


def _px(k, f, fn):
    return lambda x, p, k=k, f=f, fn=fn: p[k] * fn(x[:, f])


# ----------------------------------------------------------------------------- ChebyshevPQC
def chebyshev_pqc_num_parameters(n, num_layers=1, closed=True):
    num = 2 * n + n * num_layers
    if n > 2 and closed:
        num += n * num_layers
    else:
        num += (n - 1) * num_layers
    return num


def _chebyshev_pqc_bounds(n, num_layers, closed, alpha):
    b = []
    b += [[-np.pi, np.pi]] * n
    for _ in range(num_layers):
        b += [[0.0, alpha]] * n
        b += [[-2.0 * np.pi, 2.0 * np.pi]] * len(range(0, n, 2))
        if n > 2:
            istop = n if closed else n - 1
            b += [[-2.0 * np.pi, 2.0 * np.pi]] * len(range(1, istop, 2))
    b += [[-np.pi, np.pi]] * n
    return np.array(b, dtype=np.float64).reshape(-1, 2)


def _cheb_pqc_indices(n, num_layers, closed):
    out, off = [], n
    for _ in range(num_layers):
        out.append(list(range(off, off + n)))
        off += n
        off += len(range(0, n, 2))
        if n > 2:
            istop = n if closed else n - 1
            off += len(range(1, istop, 2))
    return out


def chebyshev_pqc_initial_parameters(n, num_layers, num_features, seed=0, closed=True, alpha=4.0):
    bounds = _chebyshev_pqc_bounds(n, num_layers, closed, alpha)
    nparam = chebyshev_pqc_num_parameters(n, num_layers, closed)
    bounds = bounds[:nparam]
    r = np.random.RandomState(seed)
    param = r.uniform(low=bounds[:, 0], high=bounds[:, 1])
    fpq = int(np.ceil(n / num_features))
    lin = np.linspace(0.01, alpha, fpq)
    for layer_idx in _cheb_pqc_indices(n, num_layers, closed):
        for i, ii in enumerate(layer_idx):
            param[ii] = lin[i % fpq]
    return param


def chebyshev_pqc(n, num_layers, num_features, closed=True, entangling_gate="crz",
                  nonlinearity="arccos", num_params=None):
    fn = _nonlin(nonlinearity)
    if entangling_gate not in ("crz", "rzz"):
        raise ValueError("Unknown entangling gate")
    d = num_features
    npar = chebyshev_pqc_num_parameters(n, num_layers, closed) if num_params is None else num_params
    gates, k, f = [], 0, 0
    for i in range(n):
        gates.append(Gate("ry", (i,), (_p(k % npar),)))
        k += 1
    for _ in range(num_layers):
        for i in range(n):
            gates.append(Gate("rx", (i,), (_px(k % npar, f % d, fn),)))
            k += 1
            f += 1
        for i in range(0, n + closed - 1, 2):
            gates.append(Gate(entangling_gate, (i, (i + 1) % n), (_p(k % npar),)))
            k += 1
        if n > 2:
            for i in range(1, n + closed - 1, 2):
                gates.append(Gate(entangling_gate, (i, (i + 1) % n), (_p(k % npar),)))
                k += 1
    for i in range(n):
        gates.append(Gate("ry", (i,), (_p(k % npar),)))
        k += 1
    return gates


# ----------------------------------------------------------------------------- Hubregtsen
def hubregtsen_num_parameters(n, num_layers=1, closed=True):
    num = n * num_layers
    if n > 2:
        num += (n if closed else n - 1) * num_layers
    return num


def hubregtsen_initial_parameters(n, num_layers, seed=0, closed=True):
    b = []
    for _ in range(num_layers):
        b += [[-np.pi, np.pi]] * n
        if n > 2:
            b += [[-2.0 * np.pi, 2.0 * np.pi]] * (n if closed else n - 1)
    b = np.array(b, dtype=np.float64).reshape(-1, 2)
    r = np.random.RandomState(seed)
    return r.uniform(low=b[:, 0], high=b[:, 1])


def hubregtsen(n, num_layers, num_features, closed=True, final_encoding=False):
    d = num_features
    npar = hubregtsen_num_parameters(n, num_layers, closed)
    gates, k = [], 0
    for i in range(n):
        gates.append(Gate("h", (i,), ()))
    loops = int(np.ceil(d / n))
    for _ in range(num_layers):
        for i in range(loops * n):
            name = "rz" if (i // n) % 2 == 0 else "rx"
            gates.append(Gate(name, (i % n,), (_x(i % d),)))
        for i in range(n):
            gates.append(Gate("ry", (i,), (_p(k % npar),)))
            k += 1
        if n > 2:
            for i in range(n if closed else n - 1):
                gates.append(Gate("crz", (i, (i + 1) % n), (_p(k % npar),)))
                k += 1
    if final_encoding:
        for i in range(loops * n):
            # reference quirk: ceil(i/n) here, i//n above (hubregtsen_encoding_circuit.py:182-188)
            name = "rz" if int(np.ceil(i / n)) % 2 == 0 else "rx"
            gates.append(Gate(name, (i % n,), (_x(i % d),)))
    return gates


# ----------------------------------------------------------------------------- ChebyshevRx
def chebyshev_rx_num_parameters(n, num_layers=1):
    return 2 * n * num_layers


def chebyshev_rx_initial_parameters(n, num_layers, num_features, seed=0, alpha=4.0):
    b = []
    for _ in range(num_layers):
        b += [[0.0, alpha]] * n
        b += [[-np.pi, np.pi]] * n
    b = np.array(b, dtype=np.float64).reshape(-1, 2)
    r = np.random.RandomState(seed)
    param = r.uniform(low=b[:, 0], high=b[:, 1])
    fpq = int(np.ceil(n / num_features))
    lin = np.linspace(0.01, alpha, fpq)
    off = 0
    for _ in range(num_layers):
        for i in range(n):
            param[off + i] = lin[i % fpq]
        off += 2 * n
    return param


def chebyshev_rx(n, num_layers, num_features, closed=False, nonlinearity="arccos"):
    fn = _nonlin(nonlinearity)
    d = num_features
    npar = chebyshev_rx_num_parameters(n, num_layers)
    gates, k, f = [], 0, 0
    for _ in range(num_layers):
        for i in range(n):
            gates.append(Gate("rx", (i,), (_px(k % npar, f % d, fn),)))
            k += 1
            f += 1
        for i in range(n):
            gates.append(Gate("rx", (i,), (_p(k % npar),)))
            k += 1
        for i in range(0, n + closed - 1, 2):
            gates.append(Gate("cx", (i, (i + 1) % n), ()))
        if n > 2:
            for i in range(1, n + closed - 1, 2):
                gates.append(Gate("cx", (i, (i + 1) % n), ()))
    return gates


# ----------------------------------------------------------------------------- Layered (test circuits)
def layered_ry_rx(n, num_features, first="ry"):
    """``LayeredEncodingCircuit(n).Ry("p") / .Rx("p")`` followed by ``.Rx("x")``.

    One gate per qubit and layer; every parameter use takes a fresh parameter, features
    wrap modulo d (layered_encoding_circuit.py:238-257).  Used by the reference's closed-form
    kernel tests (tests/kernel/lowlevel_kernel/test_fidelity_quantum_kernel.py:73-99,
    test_projected_quantum_kernel.py:58-71).
    """
    gates = []
    for i in range(n):
        gates.append(Gate(first, (i,), (_p(i),)))
    for i in range(n):
        gates.append(Gate("rx", (i,), (_x(i % num_features),)))
    return gates


# ----------------------------------------------------------------------------- HighDim
def highdim(n, num_features, num_layers=None, cycling=True, cycling_type="saw",
            layer_type="rows", entangling_gate="cx"):
    """HighDimEncodingCircuit restated for the cx entangler (iswap variant not on the path)."""
    if entangling_gate != "cx":
        raise NotImplementedError("only the cx entangler is restated")
    d = num_features
    if num_layers is None:
        num_layers = max(int(d / (n * 3)), 2)
    gates = [Gate("h", (i,), ()) for i in range(n)]
    rows = layer_type == "rows"
    off = 0
    for layer in range(num_layers):
        if layer != 0:
            for i in range(0, n - 1, 2):
                gates.append(Gate("cx", (i, i + 1), ()))
            for i in range(1, n - 1, 2):
                gates.append(Gate("cx", (i, i + 1), ()))
        for i in range(3 * n):
            iq = int(i / 3) if rows else i % n
            ii = off + i
            if cycling:
                if cycling_type == "saw":
                    ii = ii % d
                else:
                    itest = ii % max(d + d - 2, 1)
                    ii = d + d - 2 - itest if itest >= d else itest
            if iq >= n or ii >= d:
                break
            if rows:
                name = "ry" if i % 3 == 1 else "rz"
            else:
                name = "ry" if int(i / n) == 1 else "rz"
            gates.append(Gate(name, (iq,), (_x(ii),)))
        off += n * 3
        if cycling is False and off >= d:
            off = 0
    return gates
