cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so
SH=27:384:384,27:256:256,1:416:384
for cfg in "0 9" "8 9" "0 3" "0 2" "0 9" "8 9"; do
  set -- $cfg
  echo "== DEBUG=$1 STAGES=$2"
  B2ME_TC_DEBUG=$1 B2ME_TC_STAGES=$2 timeout 300 python tools/conv_probe.py --frames 8 --shapes $SH 2>&1 | tail -3
done
