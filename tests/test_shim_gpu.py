"""GPU tier: the drop-in shim (integration/soap3dp_b200_shim.cpp) EXECUTED through the reference's own declarations.
oracle/_ref/shim_check is the shim object linked with a driver written against SOAP3-dp's unmodified headers (Soap3Index,
BWT, DPParameters, GPUINDEXUpload, perform_round1_alignment, perform_round2_alignment, SemiGlobalAligner) and libsoap3dp_b200.so; it runs the case
oracle/make_shim_case.py wrote and compares every answer word, score, hit location, tie count and traced pattern with
the oracle's.  Built by oracle/build_ref.sh where /root/reference exists; the binary and its case travel to the GPU box."""
import os
import subprocess

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "shim_check")
CASE = os.path.join(ROOT, "oracle", "_ref", "shim_case")


@pytest.mark.skipif(not (os.path.exists(BIN) and os.path.exists(os.path.join(CASE, "meta.txt"))), reason="oracle/_ref/shim_check not built")
def test_the_shim_runs_through_the_reference_declarations_and_matches_the_oracle():
    r = subprocess.run([BIN, CASE], capture_output=True, text=True, timeout=300)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0, r.stdout + r.stderr
    checks = [l for l in lines if not l.startswith("INFO")]          # INFO: the timing line of the second, page-locked round-1 call
    assert len(checks) >= 10 and all(l.startswith("PASS") for l in checks), r.stdout
    assert any("straight-line kernel and check-and-extend" in l for l in lines if l.startswith("INFO")), r.stdout
    assert "drop-in shim executed through the reference's declarations" in lines[-1]
