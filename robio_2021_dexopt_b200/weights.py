"""Policy weight files, byte-compatible with dexopt's `weights-<env>-<solver>.dat`.

Reference: NeuralNetwork::saveWeights / loadWeights (dexopt/src/neural.h:375-399) walk the layers and every
DenseLayer::serialize writes its bias tensor, then its weight tensor (neural.h:193-201) as raw `Var<double>` arrays
(one double each, tractor/include/tractor/core/var.h:12-13); the weight tensor has shape (inputs, units), row-major
(dexopt/src/tensor.h:16-30, neural.h:141-147).  dexopt writes the file after training and reads it back for
`test` (dexopt/src/dexopt.cpp:47,268-282).

The optimisation vector x of this package orders each layer the way the reference registers its variables
(neural.h:153-157: weights, then bias), so a file is x with the two blocks of every layer swapped.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

Shape = Tuple[int, int]  # (inputs, units) of a dense layer


def policy_shapes(n_controlled: int, n_end_effectors: int, hidden: int, contact_dims: int = 9) -> List[Shape]:
    """dense-layer shapes of the grasp / turn policies (dexenv_grasp.h:128-146, dexenv_turn.h:102-105):
    observation = joints + 12 object features, output = joint commands + contact parameters per end effector"""
    n_in, n_out = n_controlled + 12, n_controlled + n_end_effectors * contact_dims
    return [(n_in, hidden), (hidden, n_out)] if hidden > 0 else [(n_in, n_out)]


def parameter_count(shapes: Iterable[Shape]) -> int:
    return sum(i * o + o for i, o in shapes)


def to_file_bytes(x: np.ndarray, shapes: Sequence[Shape]) -> bytes:
    x = np.ascontiguousarray(x, dtype="<f8")
    if x.size != parameter_count(shapes):
        raise ValueError("x has %d elements, the layers need %d" % (x.size, parameter_count(shapes)))
    out, off = [], 0
    for n_in, n_out in shapes:
        w = x[off:off + n_in * n_out]
        b = x[off + n_in * n_out:off + n_in * n_out + n_out]
        out += [b.tobytes(), w.tobytes()]
        off += n_in * n_out + n_out
    return b"".join(out)


def from_file_bytes(data: bytes, shapes: Sequence[Shape]) -> np.ndarray:
    need = 8 * parameter_count(shapes)
    if len(data) < need:
        # std::istream::read leaves the rest of the weights at their initial values; that is never what a caller of
        # this package wants
        raise ValueError("weight file has %d bytes, the layers need %d" % (len(data), need))
    raw = np.frombuffer(data[:need], dtype="<f8")
    out, off = [], 0
    for n_in, n_out in shapes:
        b = raw[off:off + n_out]
        w = raw[off + n_out:off + n_out + n_in * n_out]
        out += [w, b]
        off += n_in * n_out + n_out
    return np.concatenate(out)


def save_weights(path: str, x: np.ndarray, shapes: Sequence[Shape]) -> None:
    with open(path, "wb") as f:
        f.write(to_file_bytes(x, shapes))


def load_weights(path: str, shapes: Sequence[Shape]) -> np.ndarray:
    with open(path, "rb") as f:
        return from_file_bytes(f.read(), shapes)
