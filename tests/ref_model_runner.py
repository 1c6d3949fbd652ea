"""Run the UNCHANGED reference model files (model/robotnet_segmentation.py, robotnet_vote.py,
robotnet_encode.py, robotnet.py -> model/backbone/minkunet.py, resnet.py) on top of one of this repo's
MinkowskiEngine implementations and compare them with the host-side mirror (b200calib/models.py).

Separate process because utils/config.py parses sys.argv and opens its log file at import time
(SURVEY.md §5). Usage: python tests/ref_model_runner.py {oracle|cuda-construct}
"""
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "markerless-robot-camera-calibration_b200")
REF = "/root/reference"
impl = sys.argv[1] if len(sys.argv) > 1 else "oracle"

tmp = tempfile.mkdtemp(prefix="b2me_ref_")
sys.argv = ["x", "--config", os.path.join(REF, "config", "default.yaml"), "--log_path", os.path.join(tmp, "log.log"),
            "--exp_path", os.path.join(tmp, "exp")]
ipdb = types.ModuleType("ipdb")
ipdb.set_trace = lambda *a, **k: None
sys.modules["ipdb"] = ipdb
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)

import torch  # noqa: E402

if impl == "oracle":
    import oracle.MinkowskiEngine as ME
    import oracle.MinkowskiEngine.modules.resnet_block as rb
    import oracle.MinkowskiEngine.utils as mu
    import oracle.MinkowskiEngine.MinkowskiOps as mo
    sys.modules["MinkowskiEngine"] = ME
    sys.modules["MinkowskiEngine.modules"] = ME.modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = rb
    sys.modules["MinkowskiEngine.utils"] = mu
    sys.modules["MinkowskiEngine.MinkowskiOps"] = mo
else:
    import MinkowskiEngine as ME  # the CUDA-backed package: construction and state dicts work without a GPU

sys.path.insert(0, REF)
from model.robotnet_segmentation import RobotNetSegmentation as RefSeg  # noqa: E402
from model.robotnet_vote import RobotNetVote as RefVote  # noqa: E402
from model.robotnet_encode import RobotNetEncode as RefEnc  # noqa: E402
from model.robotnet import RobotNet as RefRobotNet  # noqa: E402
from model.backbone.aliveunet import AliveUNet as RefAlive, AliveUNetBase  # noqa: E402  (default.yaml: Bottleneck)
from MinkowskiEngine.modules.resnet_block import BasicBlock as rb_basic  # noqa: E402
from b200calib.models import make_models, randomize_bn_stats  # noqa: E402

M = make_models(ME)


class RefAliveBasic(AliveUNetBase):
    """the reference's base class with the BasicBlock configuration (STRUCTURE.bottleneck false, block_reps 1): the
    Bottleneck configuration of default.yaml constructs but its forward cannot run (channel arithmetic of
    aliveunet.py:123 only matches the concatenation for expansion 1)."""
    BLOCK = rb_basic
    PLANES = tuple(i * 32 for i in (list(range(1, 8)) + list(range(7, 0, -1))))
    LAYERS = (1,) * 14


def same_keys(a, b, name):
    ka = {k: tuple(v.shape) for k, v in a.state_dict().items()}
    kb = {k: tuple(v.shape) for k, v in b.state_dict().items()}
    assert ka == kb, f"{name}: state-dict mismatch {set(ka) ^ set(kb)}"
    return len(ka)


pairs = [("segmentation", RefSeg(3, num_classes=3), M.RobotNetSegmentation(3, num_classes=3)),
         ("vote", RefVote(3), M.RobotNetVote(3)),
         ("encode", RefEnc(3, 7), M.RobotNetEncode(3, 7)),
         ("robotnet", RefRobotNet(3, 7), M.RobotNet(3, 7)),
         ("aliveunet-bottleneck", RefAlive(3, 7), M.AliveUNet(3, 7, m=32, block_reps=2, bottleneck=True)),
         ("aliveunet", RefAliveBasic(3, 7), M.AliveUNet(3, 7, m=32, block_reps=1, bottleneck=False))]
for name, ref, mine in pairs:
    n = same_keys(ref, mine, name)
    print(f"{name}: {n} state-dict entries identical (keys + shapes)")

if impl == "oracle":
    torch.manual_seed(13)
    pts = torch.rand(6000, 3) * torch.tensor([50.0, 50.0, 4.0])
    feats = torch.rand(6000, 3) - 0.5
    for name, ref, mine in pairs:
        if name == "aliveunet-bottleneck":
            continue  # constructs (state dict compared above) but the reference's own forward cannot run
        randomize_bn_stats(ref)
        mine.load_state_dict(ref.state_dict())
        ref.eval(), mine.eval()
        with torch.no_grad():
            outs = []
            for net in (ref, mine):
                tf = ME.TensorField(features=feats, coordinates=ME.utils.batched_coordinates([pts], dtype=torch.float32),
                                    quantization_mode=ME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                                    minkowski_algorithm=ME.MinkowskiAlgorithm.SPEED_OPTIMIZED)
                o = net(tf.sparse())
                outs.append(o.slice(tf).F if hasattr(o, "slice") else o)
        assert torch.equal(outs[0], outs[1]), f"{name}: mirror output differs from the unchanged reference model"
        print(f"{name}: unchanged reference model == mirror on the same ME implementation, out {tuple(outs[0].shape)}")
print("OK")
