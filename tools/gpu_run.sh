set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so
for st in 2 3 4; do
B2ME_TC_STAGES=$st timeout 300 python tools/conv_probe.py --frames 8 --shapes 27:384:384 > gpurun_out/probe_st$st.log 2>&1; grep -E "^---|rank0 mma|rank0 producer" gpurun_out/probe_st$st.log
done
unset B2ME_LIB_PATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_tc_r01d -f python tools/conv_probe.py --frames 8 --shapes 27:384:384 --reps 2 > gpurun_out/ncu_full_r01d.log 2>&1; echo "ncu rc=$?"
