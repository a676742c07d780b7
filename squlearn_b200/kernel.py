"""Quantum kernels on the b200 backend: FidelityKernel and ProjectedQuantumKernel.

FidelityKernel mirrors src/squlearn/kernel/lowlevel_kernel/fidelity_kernel.py:87-232 (constructor,
``evaluate``, depolarising-noise mitigation :264-330, regularisation :228-231) on top of the
statevector path of fidelity_kernel_statevector.py:284-387: both Gram operands are simulated in
one batched engine call each and the O(N^2) Python pair loop is replaced by the FP64 tensor-core
Gram kernel with the modulus-square, symmetric mirroring and diagonal policy fused in.

ProjectedQuantumKernel mirrors projected_quantum_kernel.py:420-522, :861-973 for the Gaussian
outer kernel (sklearn outer kernels run on the host on the (N, 3n) feature matrix, as in the
reference).

KernelMatrixBase surface: kernel_matrix_base.py:64-176.
"""

from typing import Optional, Union

import numpy as np

from .data_preprocessing import adjust_features, convert_to_float64, extract_num_features
from .executor import B200Circuit, Executor, b200_evaluate_statevector
from .observables import ObservableBase, SinglePauli
from .qnn import LowLevelQNN


# ------------------------------------------------------------------------- regularisation
def thresholding_regularization(gram_matrix):
    """regularization.py:5-27: clip negative eigenvalues to zero."""
    import scipy.linalg

    evals, evecs = scipy.linalg.eig(gram_matrix)
    return np.real(evecs @ np.diag(evals.clip(min=0)) @ evecs.T)


def tikhonov_regularization(gram_matrix):
    """regularization.py:31-46: shift the spectrum by its most negative eigenvalue."""
    import scipy.linalg

    evals = scipy.linalg.eigvals(gram_matrix)
    shift = np.min(np.real(evals))
    if shift < 0:
        gram_matrix -= (shift - 1e-14) * np.identity(gram_matrix.shape[0])
    return gram_matrix


def device_kernel(kernel, x, y=None):
    """The kernel matrix of ANY kernel object as a float64 torch tensor on the device the kernel computes on:
    ``evaluate_device`` where the kernel has one (the matrix never leaves HBM), otherwise its host ``evaluate``
    (every reference kernel offers that: kernel_matrix_base.py:92-105) moved to the engine's device."""
    import torch

    if hasattr(kernel, "evaluate_device"):
        return kernel.evaluate_device(x, y).to(torch.float64)
    k = torch.as_tensor(np.asarray(kernel.evaluate(x, y), dtype=np.float64))
    dev = getattr(getattr(getattr(kernel, "_executor", None), "engine", None), "device", None)
    if dev is not None and str(dev) != "cpu":
        k = k.to(dev)
    return k


class KernelMatrixBase:
    def __init__(self, encoding_circuit, executor: Executor, initial_parameters=None,
                 parameter_seed: Union[int, None] = 0, regularization: Union[str, None] = None):
        if executor.quantum_framework != "b200":
            raise RuntimeError("squlearn_b200 kernels need Executor('b200')")
        self._encoding_circuit = encoding_circuit
        self._executor = executor
        self._parameters = None if initial_parameters is None else np.asarray(initial_parameters,
                                                                               dtype=np.float64)
        self._initial_parameters = initial_parameters
        self._parameter_seed = parameter_seed
        self._regularization = regularization
        self._num_features = None
        self._is_initialized = False

    encoding_circuit = property(lambda self: self._encoding_circuit)
    num_qubits = property(lambda self: self._encoding_circuit.num_qubits)
    num_features = property(lambda self: self._num_features)
    parameters = property(lambda self: self._parameters)
    parameter_bounds = property(lambda self: self._encoding_circuit.parameter_bounds)
    feature_bounds = property(lambda self: self._encoding_circuit.feature_bounds)

    @property
    def num_parameters(self) -> int:
        return self._encoding_circuit.num_parameters

    @property
    def is_trainable(self) -> bool:
        return self.num_parameters > 0

    def assign_parameters(self, parameters):
        self._parameters = np.asarray(parameters, dtype=np.float64)

    def evaluate_with_parameters(self, x, y, parameters):
        self.assign_parameters(parameters)
        return self.evaluate(x, y)

    def evaluate_pairwise(self, x, y=None):
        """kernel_matrix_base.py:127-140: single kernel value k(x, y)."""
        if y is None:
            y = x
        return self.evaluate(np.array([x]), np.array([y]))[0, 0]

    def _initialize_kernel(self, num_features: int) -> None:
        self._num_features = num_features
        if self._parameters is None:
            self._parameters = self._encoding_circuit.generate_initial_parameters(
                seed=self._parameter_seed, num_features=num_features)

    def _regularize_matrix(self, matrix):
        if self._regularization == "thresholding":
            return thresholding_regularization(matrix)
        if self._regularization == "tikhonov":
            return tikhonov_regularization(matrix)
        raise AttributeError("Regularization technique is unknown, please use either "
                             "'thresholding' or 'tikhonov'.")

    def _regularize_device(self, k):
        """The two regularisations of regularization.py:5-46 on a device tensor.  The kernel matrix is symmetric,
        so the eigen-decomposition is the symmetric one (device ``eigh``; the reference calls scipy's general
        ``eig`` and keeps the real part, which agrees to rounding for a symmetric matrix)."""
        import torch

        if self._regularization == "thresholding":
            evals, evecs = torch.linalg.eigh(0.5 * (k + k.T))
            return (evecs * evals.clamp(min=0.0)) @ evecs.T
        if self._regularization == "tikhonov":
            shift = torch.linalg.eigvalsh(0.5 * (k + k.T)).min()
            if float(shift) < 0:
                k = k - (shift - 1e-14) * torch.eye(k.shape[0], dtype=k.dtype, device=k.device)
            return k
        raise AttributeError("Regularization technique is unknown, please use either "
                             "'thresholding' or 'tikhonov'.")

    def evaluate(self, x, y=None):
        raise NotImplementedError


class FidelityKernel(KernelMatrixBase):
    """K_ij = |<psi(x_i)|psi(y_j)>|^2 evaluated on the B200 engine
    (kernel/lowlevel_kernel/fidelity_kernel.py:353-415 -> fidelity_kernel_statevector.py:284-387).

    Arithmetic of the Gram: small problems run FP64 DMMA; from 2^12 amplitudes and 1024 x 512 entries on the engine's
    ``precision="auto"`` policy uses the exact split-integer emulation on the INT8 tensor cores (worst-case
    |dK| <= 2 sqrt(2 * 2^n) 2^-B < 1e-10, measured ~1e-13: DESIGN.md 4.2).  ``B200SQ_GRAM_PRECISION=fp64`` keeps FP64
    DMMA everywhere."""

    def __init__(self, encoding_circuit, executor: Executor,
                 evaluate_duplicates: str = "off_diagonal", mit_depol_noise: Union[str, None] = None,
                 initial_parameters=None, parameter_seed: Union[int, None] = 0,
                 regularization: Union[str, None] = None, use_expectation: bool = False,
                 caching: bool = False):
        super().__init__(encoding_circuit, executor, initial_parameters, parameter_seed,
                         regularization)
        self._evaluate_duplicates = evaluate_duplicates.lower() if evaluate_duplicates else "none"
        if self._evaluate_duplicates not in ("all", "off_diagonal", "none"):
            raise ValueError(f"Unknown evaluate_duplicates option: {evaluate_duplicates}")
        if mit_depol_noise is not None and mit_depol_noise not in ("msplit", "mmean"):
            raise ValueError("Only msplit or mmean mitigation is available")
        if use_expectation:
            raise NotImplementedError(
                "use_expectation=True (FidelityKernelExpectationValue) is not provided by the "
                "b200 backend; the statevector path computes the same kernel")
        self._mit_depol_noise = mit_depol_noise
        self._use_expectation = use_expectation
        self._caching = caching
        self._circuit = None
        self.last_states = None  # device tensor of the x states of the last evaluate()

    def _initialize_kernel(self, num_features: int) -> None:
        if self._is_initialized and num_features == self._num_features:
            return
        super()._initialize_kernel(num_features)
        self._circuit = B200Circuit(self._encoding_circuit.get_program(num_features))
        self._is_initialized = True

    def states(self, x, as_tensor: bool = True):
        """Statevectors of the data points (N, 2^n), little-endian, on the device by default."""
        x = convert_to_float64(x)
        self._initialize_kernel(extract_num_features(x))
        if self.num_parameters > 0 and self._parameters is None:
            raise ValueError(
                "Parameters have to been set with assign_parameters or as initial parameters!")
        return self._executor.b200_execute(b200_evaluate_statevector, self._circuit,
                                           param=self._parameters, x=x, as_tensor=as_tensor)

    def evaluate_device(self, x, y=None):
        """The kernel matrix as a float64 device tensor (no device->host copy), with everything ``evaluate``
        applies — the duplicates policy, depolarising-noise mitigation (fidelity_kernel.py:264-330) and the
        regularisation (:228-231) — done on the device, so that consumers that keep K in HBM (QKRR, QGPR,
        kernel losses) see exactly what the reference's ``quantum_kernel.evaluate`` returns."""
        import torch

        num_features = extract_num_features(x)
        x = adjust_features(convert_to_float64(x), num_features)[0]
        y = x if y is None else adjust_features(convert_to_float64(y), num_features)[0]
        self._initialize_kernel(num_features)
        symmetric = np.array_equal(x, y)
        eng = self._executor.engine
        shots = self._executor.shots
        self._executor.set_shots(None)  # fidelity_kernel_statevector.py:299-300
        try:
            x_sv = self.states(x)
            self.last_states = x_sv
            if symmetric:
                # "off_diagonal"/"none": diagonal := 1.0; "all": the reference never computes the
                # symmetric diagonal and leaves it 0.0 (fidelity_kernel_statevector.py:358-383)
                diag = "zero" if self._evaluate_duplicates == "all" else "one"
                k = eng.gram(x_sv, None, diagonal=diag)
            else:
                k = eng.gram(x_sv, self.states(y))
        finally:
            self._executor.set_shots(shots)
        if self._evaluate_duplicates == "none":
            # identical statevectors are short-circuited to exactly 1.0 by the reference (:365-369, 375-378);
            # identical states <=> identical feature rows: compared on the device, in row chunks
            xd = eng.to_device(x, torch.float64)
            yd = xd if symmetric else eng.to_device(y, torch.float64)
            step = max(1, (1 << 24) // max(yd.shape[0] * max(yd.shape[1], 1), 1))
            for r0 in range(0, xd.shape[0], step):
                eq = (xd[r0:r0 + step, None, :] == yd[None, :, :]).all(dim=2)
                k[r0:r0 + step][eq] = 1.0
        if self._mit_depol_noise is not None:
            print("WARNING: Advanced option. Do not use it within an squlearn.kernel workflow")
            if not symmetric:
                raise ValueError("Mitigating depolarizing noise works only for square matrices "
                                 "computed on real backend")
            inv = 2.0 ** (-self.num_qubits)
            surv = torch.sqrt((torch.diagonal(k) - inv) / (1.0 - inv))
            if self._mit_depol_noise == "msplit":
                ss = torch.outer(surv, surv)
                k = (k - inv * (1.0 - ss)) / ss
            else:
                m2 = torch.mean(surv) ** 2
                k = (k - inv * (1.0 - m2)) / m2
        if self._regularization is not None and k.shape[0] == k.shape[1]:
            k = self._regularize_device(k)
        return k

    def evaluate(self, x, y=None) -> np.ndarray:
        return self._executor.engine.to_host(self.evaluate_device(x, y))

    def evaluate_alignment(self, x, labels):
        """(K, [sum_ij K_ij y_i y_j, sum_ij K_ij^2]) of the training kernel matrix as device tensors, with the two
        sums of the kernel-target alignment (loss/target_alignment.py:75-90) reduced inside the Gram epilogue.
        Returns None when the kernel is configured with post-processing the fused sums would not see
        (regularisation, noise mitigation, the "none" duplicates policy): callers then reduce K themselves."""
        if self._regularization is not None or self._mit_depol_noise is not None or self._evaluate_duplicates == "none":
            return None
        eng = self._executor.engine
        if not hasattr(eng, "gram_with_alignment"):
            return None
        num_features = extract_num_features(x)
        x = adjust_features(convert_to_float64(x), num_features)[0]
        self._initialize_kernel(num_features)
        x_sv = self.states(x)
        self.last_states = x_sv
        diag = "zero" if self._evaluate_duplicates == "all" else "one"
        return eng.gram_with_alignment(x_sv, np.asarray(labels, dtype=np.float64), diagonal=diag)

    def get_params(self, deep: bool = True) -> dict:
        params = {"evaluate_duplicates": self._evaluate_duplicates,
                  "mit_depol_noise": self._mit_depol_noise, "regularization": self._regularization,
                  "use_expectation": self._use_expectation, "caching": self._caching,
                  "encoding_circuit": self._encoding_circuit, "executor": self._executor}
        if deep:
            params.update(self._encoding_circuit.get_params())
        return params

    def set_params(self, **params):
        enc_keys = self._encoding_circuit.get_params()
        enc_update = {k: v for k, v in params.items() if k in enc_keys}
        for k, v in params.items():
            if k in enc_keys:
                continue
            if k not in ("evaluate_duplicates", "mit_depol_noise", "regularization", "caching"):
                raise ValueError(f"Invalid parameter {k!r}.")
            if k == "evaluate_duplicates":
                v = v.lower() if v else "none"
                if v not in ("all", "off_diagonal", "none"):
                    raise ValueError(f"Unknown evaluate_duplicates option: {v}")
            if k == "mit_depol_noise" and v is not None and v not in ("msplit", "mmean"):
                raise ValueError("Only msplit or mmean mitigation is available")
            setattr(self, "_" + k, v)
        if enc_update:
            num_parameters_backup = self.num_parameters
            self._encoding_circuit.set_params(**enc_update)
            self._is_initialized = False
            if self.num_parameters != num_parameters_backup:
                self._parameters = None
        return self


class ProjectedQuantumKernel(KernelMatrixBase):
    """k(x, y) = exp(-gamma |F(x) - F(y)|^2), F = Pauli expectation values of the encoded state."""

    def __init__(self, encoding_circuit, executor: Executor, measurement="XYZ",
                 outer_kernel="gaussian", initial_parameters=None,
                 parameter_seed: Union[int, None] = 0, regularization: Union[str, None] = None,
                 caching: bool = True, **kwargs):
        super().__init__(encoding_circuit, executor, initial_parameters, parameter_seed,
                         regularization)
        self._measurement_input = measurement
        self._caching = caching
        self._outer_kernel_name = outer_kernel if isinstance(outer_kernel, str) else None
        self._outer_kernel = outer_kernel
        self._outer_kwargs = dict(kwargs)
        self._outer_kwargs.pop("num_qubits", None)
        if isinstance(outer_kernel, str):
            if outer_kernel.lower() == "gaussian":
                self.gamma = float(self._outer_kwargs.get("gamma", 1.0))
            elif outer_kernel.lower() not in ("matern", "expsinesquared", "rationalquadratic",
                                              "dotproduct", "pairwisekernel"):
                raise ValueError(f"Unknown outer kernel: {outer_kernel}")
        elif not callable(outer_kernel):
            raise ValueError("Unknown type of outer kernel: {}".format(type(outer_kernel)))
        self._set_up_measurement_operator()
        self._qnn = None

    def _set_up_measurement_operator(self):
        m = self._measurement_input
        if isinstance(m, str):
            self._measurement = []
            for m_str in m:
                if m_str not in ("X", "Y", "Z"):
                    raise ValueError("Unknown measurement operator: {}".format(m_str))
                for i in range(self.num_qubits):
                    self._measurement.append(SinglePauli(self.num_qubits, i, op_str=m_str))
        elif isinstance(m, (ObservableBase, list)):
            self._measurement = m
        else:
            raise ValueError("Unknown type of measurement: {}".format(type(m)))

    @property
    def measurement(self):
        return self._measurement_input

    @property
    def num_parameters(self) -> int:
        num = self._encoding_circuit.num_parameters
        ms = self._measurement if isinstance(self._measurement, list) else [self._measurement]
        return num + sum(o.num_parameters for o in ms)

    def _initialize_kernel(self, num_features: int) -> None:
        if self._is_initialized and num_features == self._num_features:
            return
        had_initial = self._initial_parameters is not None
        super()._initialize_kernel(num_features)
        self._qnn = LowLevelQNN(self._encoding_circuit, self._measurement, self._executor,
                                num_features, caching=self._caching)
        if not had_initial and len(self._parameters) == self._encoding_circuit.num_parameters:
            ms = self._measurement if isinstance(self._measurement, list) else None
            if ms is not None:
                for i, o in enumerate(ms):
                    self._parameters = np.concatenate(
                        (self._parameters, o.generate_initial_parameters(seed=self._parameter_seed + i + 1)))
            else:
                self._parameters = np.concatenate(
                    (self._parameters,
                     self._measurement.generate_initial_parameters(seed=self._parameter_seed)))
        if len(self._parameters) != self.num_parameters:
            raise ValueError("Number of initial parameters is wrong, expected number: {}".format(
                self.num_parameters))
        self._is_initialized = True

    def evaluate_qnn(self, x) -> np.ndarray:
        """(N, n_obs) expectation values the outer kernel works on (:483-501)."""
        x = convert_to_float64(x)
        self._initialize_kernel(extract_num_features(x))
        nq = self._qnn.num_parameters
        param, param_op = self._parameters[:nq], self._parameters[nq:]
        return self._qnn.evaluate(x, param, param_op, "f")["f"]

    def evaluate_device(self, x, y=None):
        """The kernel matrix as a float64 device tensor (Gaussian outer kernel: computed there; other outer
        kernels: the host result of ``evaluate`` moved to the device), regularised like ``evaluate``."""
        import torch

        eng = self._executor.engine
        if self._outer_kernel_name is None or self._outer_kernel_name.lower() != "gaussian":
            return eng.to_device(self.evaluate(x, y), torch.float64)
        fx = np.atleast_2d(self.evaluate_qnn(convert_to_float64(x)))
        fy = fx if y is None else np.atleast_2d(self.evaluate_qnn(convert_to_float64(y)))
        k = eng.rbf(fx, fy, self.gamma)
        if self._regularization is not None and k.shape[0] == k.shape[1]:
            k = self._regularize_device(k)
        return k

    def evaluate(self, x, y=None) -> np.ndarray:
        x = convert_to_float64(x)
        if self._outer_kernel_name is not None and self._outer_kernel_name.lower() == "gaussian":
            # RBF(length_scale = 1/sqrt(2 gamma)) == exp(-gamma |.|^2)  (:973), on the device
            return self.evaluate_device(x, y).cpu().numpy()
        fx = np.atleast_2d(self.evaluate_qnn(x))
        fy = None
        if y is not None:
            fy = np.atleast_2d(self.evaluate_qnn(convert_to_float64(y)))
        if self._outer_kernel_name is not None:
            from sklearn.gaussian_process import kernels as skk

            cls = {"matern": skk.Matern, "expsinesquared": skk.ExpSineSquared,
                   "rationalquadratic": skk.RationalQuadratic, "dotproduct": skk.DotProduct,
                   "pairwisekernel": skk.PairwiseKernel}[self._outer_kernel_name.lower()]
            k = cls(**self._outer_kwargs)(fx, fy)
        else:
            k = self._outer_kernel(fx, fy)
        if self._regularization is not None and k.shape[0] == k.shape[1]:
            k = self._regularize_matrix(k)
        return k

    def get_params(self, deep: bool = True) -> dict:
        params = {"measurement": self._measurement_input, "outer_kernel": self._outer_kernel,
                  "regularization": self._regularization, "caching": self._caching,
                  "encoding_circuit": self._encoding_circuit, "executor": self._executor}
        if getattr(self, "gamma", None) is not None:
            params["gamma"] = self.gamma
        if deep:
            params.update(self._encoding_circuit.get_params())
        return params
