mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest17_full.log 2>&1
tail -5 gpurun_out/r2_pytest17_full.log > gpurun_out/r2_pytest17.log
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k17.log 2>&1
timeout 600 python bench.py --stages --no-cpu-baseline > gpurun_out/r2_bench17.log 2>&1
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range"
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:k_icp_eval|k_icp_build_grid|k_cluster_inter|k_kabsch|k_keypoint|k_sanity|k_translation|k_spconv_stem' -c 12 -o gpurun_out/r2_prof17_pose $CMD > gpurun_out/r2_ncu17.log 2>&1
R=gpurun_out/r2_prof17_pose.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page details --csv > gpurun_out/r2_prof17_pose_details.csv 2>/dev/null
  ncu -i $R --page raw --csv > gpurun_out/r2_prof17_pose_raw.csv 2>/dev/null
  ls -la $R
fi
tail -3 gpurun_out/r2_pytest17.log; grep "^{" gpurun_out/r2_icp1k17.log | cut -c1-200; tail -c 300 gpurun_out/r2_bench17.log
