"""TEST INFRASTRUCTURE: the slice of PennyLane's tape interface a device sees (PennyLane is not in this image).
Operations carry ``name`` / ``wires`` / ``parameters`` / ``hyperparameters``; a tape has ``operations`` and
``measurements``; measurement classes are named like PennyLane's (StateMP, ExpectationMP) and observables like its
operator arithmetic (PauliX/Y/Z, Identity, Prod, SProd, Sum)."""

import numpy as np


class Op:
    def __init__(self, name, wires, parameters=(), **hyper):
        self.name, self.wires = name, list(wires)
        self.parameters = [np.asarray(p, dtype=float) for p in parameters]
        self.hyperparameters = hyper

    @property
    def batch_size(self):
        sizes = [p.shape[0] for p in self.parameters if p.ndim == 1]
        return sizes[0] if sizes else None

    def __repr__(self):
        return f"{self.name}(wires={self.wires})"


def _obs(name):
    return lambda wire: Op(name, [wire])


PauliX, PauliY, PauliZ, Identity = _obs("PauliX"), _obs("PauliY"), _obs("PauliZ"), _obs("Identity")


class Prod:
    name = "Prod"

    def __init__(self, *operands):
        self.operands = list(operands)

    def __repr__(self):
        return "Prod(" + ", ".join(map(repr, self.operands)) + ")"


class SProd:
    name = "SProd"

    def __init__(self, scalar, base):
        self.scalar, self.base = scalar, base

    def __repr__(self):
        return f"SProd({self.scalar}, {self.base!r})"


class Sum:
    name = "Sum"

    def __init__(self, *operands):
        self.operands = list(operands)

    def __repr__(self):
        return "Sum(" + ", ".join(map(repr, self.operands)) + ")"


class StateMP:
    obs = None


class ExpectationMP:
    def __init__(self, obs):
        self.obs = obs


class QuantumScript:
    def __init__(self, operations, measurements):
        self.operations, self.measurements = list(operations), list(measurements)
