"""TEST INFRASTRUCTURE: an engine stand-in that runs the product's plan + per-thread kernel code on the CPU
(tests/hostemu), so that the Python host glue of the package (b200_evaluate / b200_gradient /
LowLevelQNN.evaluate ...) can be exercised by the CPU suite exactly as it runs on the GPU engine.  Never
imported by the package; not a fallback for anything."""

import numpy as np
import torch

from squlearn_b200.engine import CompiledProgram
from squlearn_b200.observables import pauli_masks
from tests import hostemu


class EmuEngine:
    device = "cpu"

    def compile(self, program, params=None, tile_bits=0):
        return CompiledProgram(program, params, tile_bits)   # planning is host-only

    def _states(self, cprog, x, variants):
        x = np.ascontiguousarray(x, dtype=np.float64)
        psi, _ = hostemu.simulate(cprog.program, cprog.params, x, variants=variants)
        return psi

    def simulate(self, cprog, x, variants=None):
        return torch.from_numpy(self._states(cprog, x, variants))

    def simulate_expvals(self, cprog, x, terms, variants=None):
        n = cprog.program.num_qubits
        psi = self._states(cprog, x, variants)
        masks = [pauli_masks(t) for t in terms]
        out = hostemu.expvals(psi.reshape(-1, 1 << n), n, [m[0] for m in masks], [m[1] for m in masks])
        return torch.from_numpy(out.reshape(psi.shape[:-1] + (len(terms),)))


    # --- consumers of the states: plain NumPy stand-ins of the Gram / RBF kernels (the CUDA kernels have
    # their own parity tests; here only the host glue around them is under test)
    def gram(self, psi_x, psi_y=None, diagonal="one", out=None, **_):
        a = psi_x.numpy()
        if psi_y is None:
            k = np.abs(a.conj() @ a.T) ** 2
            k = np.tril(k) + np.tril(k, -1).T
            if diagonal == "one":
                np.fill_diagonal(k, 1.0)
            elif diagonal == "zero":
                np.fill_diagonal(k, 0.0)
        else:
            k = np.abs(a.conj() @ psi_y.numpy().T) ** 2
        return torch.from_numpy(k)

    def rbf(self, fx, fy, gamma):
        fx, fy = np.asarray(fx, dtype=np.float64), np.asarray(fy, dtype=np.float64)
        d2 = ((fx[:, None, :] - fy[None, :, :]) ** 2).sum(axis=2)
        return torch.from_numpy(np.exp(-gamma * d2))

    def to_host(self, t):
        return t.numpy().copy()

    def to_device(self, a, dtype=None):
        return torch.as_tensor(np.asarray(a))

    def launch_count(self):
        return 0


class EmuExecutor:
    """The slice of squlearn_b200.Executor the QNN / kernel glue uses."""
    quantum_framework = "b200"
    shots = None

    def __init__(self):
        self.engine = EmuEngine()

    def set_shots(self, shots):
        self.shots = shots

    def b200_execute(self, fn, circuit, **kwargs):
        return fn(circuit, executor=self, **kwargs)
