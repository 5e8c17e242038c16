"""GPU: the additive batch API of the header (QRBatchFactoriser) against the reference-shaped
single-matrix QRFactoriser looped over the batch, plus the properties the reference's QR tests check. The C++ program is
tests/host_harness/additive_test.cu, built by __graft_entry__.build() into build/tests/."""
import subprocess

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
BIN = REPO / "build" / "tests" / "additive_test"


def test_batch_qr_equals_the_single_matrix_factoriser_looped():
    if not BIN.exists():
        pytest.fail(f"{BIN} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    r = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "ALL PASSED" in out and "FAIL " not in out, out[-3000:]
    assert out.count("PASS ") >= 21
