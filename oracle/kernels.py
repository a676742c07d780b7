"""Fidelity / projected quantum kernels (TEST INFRASTRUCTURE — see oracle/__init__.py).

  * ``fidelity_gram_pairloop`` restates FidelityKernelStatevector.evaluate_kernel_sv
    (src/squlearn/kernel/lowlevel_kernel/fidelity_kernel_statevector.py:284-387): symmetric
    detection by ``np.array_equal(x, y)`` (:301), entry ``abs(matmul(x.conj(), y))**2`` (:303-310),
    j < i loop with mirroring (:358-371), rectangular double loop (:372-379), diagonal policy
    (:381-383) — including the quirk that ``evaluate_duplicates="all"`` leaves the symmetric
    diagonal at 0.0.  This is the reference's hot loop B and the timed CPU baseline.
  * ``fidelity_gram_zgemm`` is the same quantity through one BLAS zgemm (checker at larger N).
  * ``projected_quantum_kernel`` restates ProjectedQuantumKernel.evaluate with the Gaussian outer
    kernel (src/squlearn/kernel/lowlevel_kernel/projected_quantum_kernel.py:861-877, :945-973):
    K = RBF(length_scale = 1/sqrt(2 gamma))(F_x, F_y) = exp(-gamma |F_i - F_j|^2).
"""

import numpy as np

from .statevector import simulate
from .observables import pauli_expectation, xyz_measurement_strings


def fidelity_gram_pairloop(x_sv, y_sv, symmetric, evaluate_duplicates="off_diagonal"):
    x_sv = np.asarray(x_sv)
    y_sv = np.asarray(y_sv)
    kernel_matrix = np.zeros((x_sv.shape[0], y_sv.shape[0]))
    if symmetric:
        for i in range(len(x_sv)):
            for j in range(i):
                if np.array_equal(x_sv[i], y_sv[j]):
                    if evaluate_duplicates == "none":
                        kernel_matrix[i, j] = 1.0
                        kernel_matrix[j, i] = 1.0
                        continue
                kernel_matrix[i, j] = np.abs(np.matmul(x_sv[i].conj(), y_sv[j])) ** 2
                kernel_matrix[j, i] = kernel_matrix[i, j]
    else:
        for i, x_ in enumerate(x_sv):
            for j, y_ in enumerate(y_sv):
                if np.array_equal(x_, y_):
                    if evaluate_duplicates == "none":
                        kernel_matrix[i, j] = 1.0
                        continue
                kernel_matrix[i, j] = np.abs(np.matmul(x_.conj(), y_)) ** 2
    if evaluate_duplicates in ["none", "off_diagonal"] and symmetric:
        for i in range(x_sv.shape[0]):
            kernel_matrix[i, i] = 1.0
    return kernel_matrix


def fidelity_gram_zgemm(x_sv, y_sv, symmetric, evaluate_duplicates="off_diagonal"):
    k = np.abs(np.asarray(x_sv).conj() @ np.asarray(y_sv).T) ** 2
    if symmetric:
        k = 0.5 * (k + k.T)
        if evaluate_duplicates in ("none", "off_diagonal"):
            np.fill_diagonal(k, 1.0)
        else:
            np.fill_diagonal(k, 0.0)
    return k


def fidelity_kernel(gates, n_qubits, x, y=None, p=None, evaluate_duplicates="off_diagonal",
                    pairloop=True):
    x = np.asarray(x, dtype=np.float64)
    y = x if y is None else np.asarray(y, dtype=np.float64)
    symmetric = np.array_equal(x, y)
    x_sv = simulate(gates, n_qubits, x, p)
    y_sv = x_sv if symmetric else simulate(gates, n_qubits, y, p)
    fn = fidelity_gram_pairloop if pairloop else fidelity_gram_zgemm
    return fn(x_sv, y_sv, symmetric, evaluate_duplicates)


def projected_kernel_features(gates, n_qubits, x, p=None, measurement="XYZ"):
    sv = simulate(gates, n_qubits, x, p)
    strings = xyz_measurement_strings(n_qubits, measurement)
    return np.stack([pauli_expectation(sv, s) for s in strings], axis=1)


def gaussian_outer_kernel(fx, fy, gamma=1.0):
    fx = np.asarray(fx, dtype=np.float64)
    fy = np.asarray(fy, dtype=np.float64)
    d2 = np.sum((fx[:, None, :] - fy[None, :, :]) ** 2, axis=2)
    return np.exp(-gamma * d2)


def projected_quantum_kernel(gates, n_qubits, x, y=None, p=None, gamma=1.0, measurement="XYZ"):
    fx = projected_kernel_features(gates, n_qubits, x, p, measurement)
    fy = fx if y is None else projected_kernel_features(gates, n_qubits, y, p, measurement)
    return gaussian_outer_kernel(fx, fy, gamma)
