"""squlearn_b200 — B200-native batched statevector engine for sQUlearn's data-parallel hot path.

Host side (Python) mirrors the reference's interface for this path; all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/b200sq.h`` (``libb200sq.so``).
There is no CPU fallback: constructing ``Executor("b200")`` without a CUDA device raises.
"""

from .circuits import (
    ChebyshevPQC,
    ChebyshevRx,
    EncodingCircuitBase,
    HighDimEncodingCircuit,
    HubregtsenEncodingCircuit,
    LayeredEncodingCircuit,
)
from .executor import (
    B200Circuit,
    Executor,
    b200_evaluate,
    b200_evaluate_statevector,
    b200_gradient,
    b200_operator_gradient,
)
from .kernel import FidelityKernel, ProjectedQuantumKernel
from .kernel_loss import NLL, KernelOptimizer, ScipyMinimize, TargetAlignment
from .observables import CustomObservable, SinglePauli, SummedPaulis
from .program import Angle, GateOp, Program
from .qgpr import QGPR
from .qkrr import QKRR
from .qnn import LowLevelQNN
from .qnn_training import (SLSQP, LBFGSB, Adam, MeanSquaredError, QNNLossBase, QNNRegressor, SquaredLoss,
                           VarianceLoss, train)
from .qsvc import QSVC
from .qrc import QRCRegressor, QRCClassifier, random_pauli_list

__version__ = "0.1.0"

__all__ = [
    "Executor", "B200Circuit", "FidelityKernel", "ProjectedQuantumKernel", "LowLevelQNN",
    "ChebyshevPQC", "ChebyshevRx", "HubregtsenEncodingCircuit", "HighDimEncodingCircuit",
    "LayeredEncodingCircuit", "EncodingCircuitBase", "SinglePauli", "SummedPaulis",
    "CustomObservable", "Program", "GateOp", "Angle", "QRCRegressor", "QRCClassifier",
    "random_pauli_list", "b200_evaluate", "b200_evaluate_statevector", "b200_gradient",
    "b200_operator_gradient", "QKRR", "QGPR", "QSVC", "QNNRegressor", "SquaredLoss", "MeanSquaredError", "VarianceLoss", "QNNLossBase", "train",
    "SLSQP", "LBFGSB", "Adam", "TargetAlignment", "NLL", "KernelOptimizer", "ScipyMinimize",
]
