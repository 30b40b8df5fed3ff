"""Phase-1 parity against the committed golden vectors (tests/golden/align_tiny.npz, produced by the
unmodified reference through tools/make_golden.py).  Needs neither /root/reference nor oracle/_ref."""
import pytest

import smoke_check


def test_golden_hostemu(hostemu):
    smoke_check.run(hostemu)


@pytest.mark.gpu
def test_golden_cuda(cuda):
    smoke_check.run(cuda)
