"""Shape contract of the public API (mirrors src/squlearn/util/data_preprocessing.py:6-154).

``adjust_features`` / ``adjust_parameters`` bring scalars / 1-D / 2-D inputs to the 2-D form and
report whether the caller passed several inputs; they raise the same ValueErrors as the reference.
"""

from typing import Tuple, Union

import numpy as np


def convert_to_float64(x) -> np.ndarray:
    if not isinstance(x, np.ndarray):
        x = np.array(x)
    if x.dtype != np.float64:
        x = np.real_if_close(x)
        if np.iscomplexobj(x):
            raise ValueError("Only real values for parameters and features are supported in sQUlearn!")
        x = np.array(x, dtype=np.float64)
    return x


def _adjust_input(x, x_length: int, allow_single_array: bool) -> Tuple[np.ndarray, bool]:
    multiple_inputs = False
    shape = np.shape(x)
    if shape == () and x_length == 1:
        xx = np.array([[x]])
    elif sum(shape) == 0 and x_length > 0:
        raise ValueError(f"Received an empty input, but expected {x_length} features.")
    elif len(shape) == 1:
        if x_length == 1:
            xx = np.array([np.array([v]) for v in x])
            multiple_inputs = (shape[0] != 1) if allow_single_array else True
        else:
            if len(x) == x_length:
                xx = np.array([x])
            else:
                raise ValueError(f"1D input has length {len(x)}, but expected {x_length} features.")
    elif len(shape) == 2:
        if shape[1] == x_length:
            xx = x
            multiple_inputs = True
        else:
            raise ValueError(
                f"2D input has shape {shape}, but second dimension must match expected feature "
                f"dimension {x_length}.")
    else:
        raise ValueError(
            f"Unsupported input shape {shape}. Expected scalar, 1D array of length {x_length}, "
            f"or 2D array with shape (n_samples, {x_length}).")
    return convert_to_float64(xx), multiple_inputs


def adjust_features(x: Union[np.ndarray, float], x_length: int) -> Tuple[np.ndarray, bool]:
    return _adjust_input(x, x_length, allow_single_array=False)


def adjust_parameters(x, x_length: int) -> Tuple[np.ndarray, bool]:
    return _adjust_input(x, x_length, allow_single_array=True)


def extract_num_features(X) -> int:
    if isinstance(X, list):
        X = np.array(X)
    if isinstance(X, np.ndarray):
        if X.ndim == 1:
            return X.shape[0]
        if X.ndim >= 2:
            return X.shape[1]
    raise TypeError("Unsupported type for X")
