"""Label parity on the BENCHED configuration (BASELINE.json configs[1]: 640x480 Kinect-shaped frames, 5 mm voxels,
MinkUNet18D, the bench's seed-13 weights): per-point arg-max labels of the CUDA path in every compute mode against
the fp32 CPU oracle, counted over ALL points, no margin mask (app/inference_engine.py:405-424,
utils/output.py:67-73). A summation order different from the oracle's BLAS makes exact ties unwinnable, so the
assertion is a bound on the measured mismatch COUNT, which the test prints; the bench JSON carries the same numbers
(`parity`)."""
import numpy as np
import pytest
import torch

import bench
import oracle.MinkowskiEngine as OME
from oracle import pipeline as OP
from gpu_util import rel_err

pytestmark = pytest.mark.gpu

# measured on B200 (see profiles/README.md); bounds = a few times the measured count
BOUNDS = {"f32": 2e-5, "tf32": 2e-3, "bf16": 3e-2}
LOGIT_TOL = {"f32": 1e-3, "tf32": 1e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def fullsize():
    import MinkowskiEngine as ME
    frames = bench.make_workload(2, 0, 640, 480)                 # the bench's own frames 0 and 1
    oseg, _, _ = bench.build_models(OME)
    cseg = bench.build_models(ME)[0].cuda()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = []
    for fr in frames:                                             # the reference runs one frame at a time
        p, c = fr[0], fr[1]
        with torch.no_grad():
            _, raw = OP.predict_segmentation(oseg, p, OP.normalize_colors(c), 200.0)
        ref.append(raw)
    return ME, frames, cseg, ref


@pytest.mark.parametrize("mode", ["f32", "tf32", "bf16"])
def test_fullsize_label_mismatch_count(fullsize, mode):
    ME, frames, cseg, ref = fullsize
    from b200calib.pipeline import segment_points, normalize_colors_
    counts = [len(f[0]) for f in frames]
    pts = torch.from_numpy(np.concatenate([f[0] for f in frames])).cuda()
    rgb = torch.from_numpy(np.concatenate([f[1] for f in frames])).cuda()
    bidx = torch.from_numpy(np.repeat(np.arange(len(frames), dtype=np.float32), counts)).cuda()
    ME.set_compute_dtype(mode)
    try:
        with torch.no_grad():
            labels, fld, out = segment_points(cseg, pts, normalize_colors_(rgb), bidx, len(frames), 200.0)
            plog = out.slice(fld).F.float().cpu()
        labels = labels.cpu().numpy()
    finally:
        ME.set_compute_dtype(torch.float32)
    ref_lab = np.concatenate([r["labels"] for r in ref])
    ref_log = torch.from_numpy(np.concatenate([r["point_logits"] for r in ref]))
    N = len(ref_lab)
    mism = int((labels != ref_lab).sum())
    err = rel_err(plog, ref_log)
    hist = np.bincount(ref_lab, minlength=3).tolist()
    print(f"[{mode}] full-size labels: {mism} of {N} points differ from the fp32 oracle ({mism / N:.2e}); "
          f"point logits rel err {err:.2e}; oracle label histogram {hist}")
    assert len(set(np.unique(ref_lab))) >= 2, "degenerate workload: the oracle predicts a single class"
    # the fused arg-max (head epilogue / b2me_linear_small + b2me_gather_labels) is the arg-max of the sliced logits
    assert np.array_equal(labels, plog.max(1)[1].numpy().astype(labels.dtype))
    assert err < LOGIT_TOL[mode]
    assert mism <= BOUNDS[mode] * N, f"{mism} mismatching labels of {N} in {mode} mode"
    # every mismatch sits where the oracle's own top-2 margin is within the logit error of this mode
    bad = np.nonzero(labels != ref_lab)[0]
    if len(bad):
        margin = np.concatenate([r["margin"] for r in ref])[bad]
        maxdiff = float((plog - ref_log).abs().max())
        assert float(margin.max()) <= 2.0 * maxdiff + 1e-12
