mkdir -p gpurun_out
P=$PWD/markerless-robot-camera-calibration_b200
V=$P/lib_variants
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest10_full.log 2>&1
tail -5 gpurun_out/r2_pytest10_full.log > gpurun_out/r2_pytest10.log
PR="timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep"
$PR --lib-b $V/libb2me_nohint.so --shapes 27:384:384,8:384:384,1:416:384,1:384:256 > gpurun_out/r2_probe10_l1.log 2>&1
$PR --lib-b $V/libb2me_nohint.so --head 3 --shapes 1:256:1024 > gpurun_out/r2_probe10_head.log 2>&1
$PR --lib-b $V/libb2me_nohint.so --level 2 --shapes 27:384:384,27:32:32 > gpurun_out/r2_probe10_l2.log 2>&1
$PR --lib-b $V/libb2me_trim.so --shapes 27:384:384,1:416:384,8:384:384 > gpurun_out/r2_probe10_l1_trim.log 2>&1
$PR --lib-b $V/libb2me_trim.so --head 3 --shapes 1:256:1024 > gpurun_out/r2_probe10_head_trim.log 2>&1
timeout 600 python bench.py --stages --crop both --conv-table gpurun_out/r2_conv_table10.json --torch-profile gpurun_out/r2_torch_profile10.txt > gpurun_out/r2_bench10.log 2>&1
B2ME_LIB_PATH=$V/libb2me_nohint.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench10_nohint.log 2>&1
timeout 300 python bench.py --dtype tf32 --steps 5 > gpurun_out/r2_bench10_tf32.log 2>&1
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k10.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384,8:384:384,1:416:384 > gpurun_out/r2_roles10_l1.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --level 2 --shapes 27:384:384 > gpurun_out/r2_roles10_l2.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --head 3 --shapes 1:256:1024 > gpurun_out/r2_roles10_head.log 2>&1
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range"
$CMD > gpurun_out/r2_plain10.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches10.csv $CMD > gpurun_out/r2_ncu10_a.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 40 -c 9 -o gpurun_out/r2_prof10_tc $CMD > gpurun_out/r2_ncu10_b.log 2>&1
R=gpurun_out/r2_prof10_tc.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page details --csv > gpurun_out/r2_prof10_tc_details.csv 2>/dev/null
  ncu -i $R --page raw --csv > gpurun_out/r2_prof10_tc_raw.csv 2>/dev/null
  for i in 0 1 2 3 4 5 6 7 8; do ncu -i $R --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/r2_prof10_tc_src$i.csv.gz; done
  rm -f $R
fi
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:k_hash_insert|k_assign_rows|k_inverse_accumulate|k_kernel_map_k3_blocks|k_kernel_map_k3$|k_block_rows|k_row_masks|k_mask_keys_rows|k_tile_masks_rows|k_spconv_stem|k_cluster_|k_icp_eval|k_color_|k_select_|k_gather_crops|k_global_pool|k_kabsch|k_translation|k_sanity|k_first_flags|k_stride_kernel_maps|k_quantize' -c 40 -o gpurun_out/r2_prof10_misc $CMD > gpurun_out/r2_ncu10_c.log 2>&1
R=gpurun_out/r2_prof10_misc.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page details --csv > gpurun_out/r2_prof10_misc_details.csv 2>/dev/null
  ncu -i $R --page raw --csv > gpurun_out/r2_prof10_misc_raw.csv 2>/dev/null
  rm -f $R
fi
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spconv_tc --csv --log-file gpurun_out/r2_tc_dram10.csv $CMD > gpurun_out/r2_ncu10_d.log 2>&1
du -sh gpurun_out; tail -3 gpurun_out/r2_pytest10.log; grep -h "median" gpurun_out/r2_probe10_*.log | grep -E "default" ; tail -c 300 gpurun_out/r2_bench10.log
