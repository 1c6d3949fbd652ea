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
    for name in ("segmentation", "vote", "encode", "robotnet", "aliveunet"):
        assert f"{name}: unchanged reference model == mirror" in out


@pytest.mark.reference
def test_reference_models_construct_on_cuda_me(built_lib):
    out = _run("cuda-construct")
    assert "segmentation: 294 state-dict entries identical" in out


@pytest.mark.reference
def test_reference_pointnet2_state_dict_matches_mirror(built_lib):
    """model/pointnet2.py (unchanged) and b200calib.pointnet2 have identical parameters under the same seed, and the
    unchanged reference file constructs on top of this repo's drop-in pointnet2_utils module."""
    code = r'''
import sys, types, torch
m = types.ModuleType("ipdb"); m.set_trace = lambda *a, **k: None; sys.modules["ipdb"] = m
sys.path.insert(0, "%s"); sys.path.insert(0, "%s")
from b200calib import pointnet2_utils as U, pointnet2 as P
sys.path.insert(0, "/root/reference")
import model  # the reference package
sys.modules["model.pointnet2_utils"] = U          # drop-in: the reference model file imports it unchanged
from model.pointnet2 import PointNet2SSG as Ref
torch.manual_seed(13); a = Ref(6, in_channels=6)
torch.manual_seed(13); b = P.PointNet2SSG(6, in_channels=6)
sa, sb = a.state_dict(), b.state_dict()
assert list(sa) == list(sb)
assert all(torch.equal(sa[k], sb[k]) for k in sa)
assert type(a.sa1).__module__.endswith("b200calib.pointnet2_utils")
print("OK", len(sa))
''' % (ROOT, os.path.join(ROOT, "markerless-robot-camera-calibration_b200"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.strip().startswith("OK")
