"""CPU: the .bt tensor file format (ref: tensor.cuh:762-816, python/gputils_api/gputils_api.py:3-52)."""
import struct

import numpy as np

import bt_format


def test_header_and_payload_order(tmp_path):
    x = bt_format.reference_b_d()              # the array the reference's python test writes (python/test/test.py:9-13)
    p = str(tmp_path / "b_d.bt")
    bt_format.write_bt(p, x)
    raw = open(p, "rb").read()
    assert struct.unpack("<QQQ", raw[:24]) == (3, 3, 2)
    payload = np.frombuffer(raw[24:], dtype=np.float64)
    # what testTensor.cu:193-202 expects: b(i, j, k) = 1 + 2j + 6i + k at i + 3*(j + 3*k)
    for i in range(3):
        for j in range(3):
            for k in range(2):
                assert payload[i + 3 * (j + 3 * k)] == 1 + 2 * j + 6 * i + k
    assert np.array_equal(bt_format.read_bt(p), x)


def test_round_trip_shapes_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    for shape in [(5,), (4, 5), (2, 4, 6)]:
        for dt in (np.float64, np.float32):
            x = rng.standard_normal(shape).astype(dt)
            p = str(tmp_path / "t.bt")
            bt_format.write_bt(p, x)
            y = bt_format.read_bt(p, dtype=dt)
            assert y.shape == shape + (1,) * (3 - len(shape))
            assert np.array_equal(y.reshape(shape), x)
