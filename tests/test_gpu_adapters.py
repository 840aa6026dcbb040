"""GPU: the C++ drop-in adapters (adapters/) side by side with the reference's own classes.

adapters/_build/test_adapters is built in the build container against /root/reference
(adapters/Makefile, via __graft_entry__.build()) and travels to the GPU box; it exits 0 only if
every ScanMatchingSummary / LoopDetectionResult field is bit-identical."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adapters", "_build",
                   "test_adapters")


@pytest.mark.skipif(not os.path.exists(BIN), reason="adapters/_build/test_adapters not built "
                    "(needs the reference tree at build time)")
def test_adapters_identical_to_reference_classes():
    p = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "ALL IDENTICAL" in p.stdout
