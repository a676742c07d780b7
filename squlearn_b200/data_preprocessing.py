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
    """Rows of width ``x_length`` from a scalar / 1-D / 2-D input, and whether the caller passed several rows.

    Shape contract (and messages) of src/squlearn/util/data_preprocessing.py:37-91: a 1-D input of width-1 data
    is a column of samples — always "several" for features, only when longer than one for parameters; a 1-D
    input of wider data is one row; an input without any extent is rejected when a width is expected."""
    arr = np.asarray(x)
    ndim, width = arr.ndim, int(x_length)
    if ndim == 0 and width == 1:
        return convert_to_float64(arr.reshape(1, 1)), False
    if width > 0 and not any(arr.shape):
        raise ValueError(f"Received an empty input, but expected {x_length} features.")
    if ndim == 1:
        if width == 1:
            return convert_to_float64(arr.reshape(-1, 1)), (arr.shape[0] != 1 or not allow_single_array)
        if arr.shape[0] != width:
            raise ValueError(f"1D input has length {arr.shape[0]}, but expected {x_length} features.")
        return convert_to_float64(arr.reshape(1, width)), False
    if ndim == 2:
        if arr.shape[1] != width:
            raise ValueError(
                f"2D input has shape {arr.shape}, but second dimension must match expected feature "
                f"dimension {x_length}.")
        return convert_to_float64(arr), True
    raise ValueError(
        f"Unsupported input shape {arr.shape}. Expected scalar, 1D array of length {x_length}, "
        f"or 2D array with shape (n_samples, {x_length}).")


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
