"""GPU: drop-in proof.  The reference's own Python layer (Signal, ProtocolAnalyzer, Modulator,
AutoInterpretation, ... — unmodified, from /root/reference or the staged copy oracle/_ref/pyref) with
urh.cythonext.{signal_functions, util, auto_interpretation} replaced by the urh_b200 shims runs the reference's own hot-path
tests: tests/test_demodulations.py, test_modulator.py, test_iq_array.py, test_protocol_analyzer.py, test_ringbuffer.py and
tests/auto_interpretation/* (79 tests with the reference's kernels)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(impl):
    return subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_tests.py"), "--impl", impl],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500, cwd=ROOT)


def test_reference_suite_passes_on_the_b200_kernels():
    out = _run("b200")
    if "REFERENCE_TESTS unavailable" in out.stdout:
        pytest.skip("reference python layer not staged (oracle/build_ref.py needs /root/reference)")
    assert "substituted urh.cythonext.signal_functions" in out.stdout
    assert "REFERENCE_TESTS impl=b200 rc=0" in out.stdout, out.stdout[-6000:]
