"""The unchanged reference model files run on this repo's MinkowskiEngine implementations and agree with the
host-side mirror used on the GPU box (authoring container only: needs /root/reference)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(impl):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_model_runner.py"), impl], capture_output=True,
                       text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.strip().endswith("OK")
    return p.stdout


@pytest.mark.reference
def test_reference_models_run_unchanged_on_oracle_me():
    out = _run("oracle")
    for name in ("segmentation", "vote", "encode", "robotnet"):
        assert f"{name}: unchanged reference model == mirror" in out


@pytest.mark.reference
def test_reference_models_construct_on_cuda_me(built_lib):
    out = _run("cuda-construct")
    assert "segmentation: 294 state-dict entries identical" in out
