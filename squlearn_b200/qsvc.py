"""Quantum support vector classification on the B200 kernels (BASELINE configs[0]).

Mirror of ``squlearn.kernel.QSVC`` (src/squlearn/kernel/qsvc.py:17-147): a scikit-learn ``SVC`` whose kernel is
the quantum kernel's ``evaluate`` (or ``"precomputed"``).  The SMO solver stays scikit-learn's (host); every
Gram block it asks for is one ``FidelityKernel.evaluate`` / ``ProjectedQuantumKernel.evaluate`` call on the GPU.
"""

from typing import Optional, Union

from sklearn.svm import SVC

from .data_preprocessing import extract_num_features
from .kernel import KernelMatrixBase


class QSVC(SVC):
    def __init__(self, quantum_kernel: Optional[Union[KernelMatrixBase, str]] = None, **kwargs) -> None:
        self._quantum_kernel = quantum_kernel
        self._kernel_params = kwargs
        superclass_params = set(self._get_param_names()) & self._kernel_params.keys()
        if isinstance(self._quantum_kernel, str) or self._quantum_kernel is None:
            kernel_argument = "precomputed"
        else:  # a kernel object: use its evaluate method (qsvc.py:83-87)
            kernel_argument = self._quantum_kernel.evaluate
        super().__init__(kernel=kernel_argument, **{key: self._kernel_params[key] for key in superclass_params})

    @classmethod
    def _get_param_names(cls):
        names = SVC._get_param_names()
        for drop in ("kernel", "gamma", "degree", "coef0"):
            names.remove(drop)
        return names

    @property
    def quantum_kernel(self):
        return self._quantum_kernel

    def fit(self, X, y, sample_weight=None):
        if not isinstance(self._quantum_kernel, str) and self._quantum_kernel is not None:
            self._quantum_kernel._initialize_kernel(num_features=extract_num_features(X))
        return super().fit(X, y, sample_weight=sample_weight)

    def get_params(self, deep: bool = True) -> dict:
        params = {"quantum_kernel": self._quantum_kernel}
        for key in self._get_param_names():
            params[key] = getattr(self, key)
        return params

    def set_params(self, **params) -> "QSVC":
        if "quantum_kernel" in params:
            self._quantum_kernel = params.pop("quantum_kernel")
            self.kernel = "precomputed" if isinstance(self._quantum_kernel, str) else self._quantum_kernel.evaluate
        for key, value in params.items():
            if key not in self._get_param_names():
                raise ValueError(f"Invalid parameter {key!r}.")
            setattr(self, key, value)
        return self
