"""weights-<env>-<solver>.dat (dexopt/src/neural.h:193-201,375-399): bias then weights per dense layer, raw doubles."""
import numpy as np
import pytest

from robio_2021_dexopt_b200 import weights


def test_layout_follows_dense_layer_serialize(tmp_path):
    shapes = weights.policy_shapes(29, 6, 64)
    assert shapes == [(41, 64), (64, 83)] and weights.parameter_count(shapes) == 8083
    x = np.arange(8083, dtype=np.float64)
    raw = np.frombuffer(weights.to_file_bytes(x, shapes), dtype="<f8")
    # layer 1: bias (x[41*64 : 41*64+64]) first, then the (41, 64) weight tensor row-major
    assert np.array_equal(raw[:64], x[41 * 64:41 * 64 + 64])
    assert np.array_equal(raw[64:64 + 41 * 64], x[:41 * 64])
    off = 41 * 64 + 64
    assert np.array_equal(raw[off:off + 83], x[off + 64 * 83:off + 64 * 83 + 83])
    assert np.array_equal(raw[off + 83:], x[off:off + 64 * 83])
    path = str(tmp_path / "weights-grasp-gd.dat")
    weights.save_weights(path, x, shapes)
    assert np.array_equal(weights.load_weights(path, shapes), x)


def test_turn_policy_is_one_linear_layer():
    assert weights.policy_shapes(24, 5, 0) == [(36, 69)]


def test_short_files_are_refused():
    shapes = [(3, 2)]
    with pytest.raises(ValueError):
        weights.from_file_bytes(b"\0" * 8 * 7, shapes)
    with pytest.raises(ValueError):
        weights.to_file_bytes(np.zeros(7), shapes)
