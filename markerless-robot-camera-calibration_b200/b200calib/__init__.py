"""b200calib — host-side mirror of the reference's inference pipeline (app/inference_engine.py,
utils/output.py, utils/transformation.py, utils/icp.py, utils/calibration.py) on top of libb2me and the
B200 MinkowskiEngine-compatible package. Batched over frames; one process per GPU."""
