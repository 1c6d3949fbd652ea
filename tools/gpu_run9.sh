mkdir -p gpurun_out
P=$PWD/markerless-robot-camera-calibration_b200
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest9_full.log 2>&1
tail -5 gpurun_out/r2_pytest9_full.log > gpurun_out/r2_pytest9.log
timeout 600 python bench.py --stages --crop both --conv-table gpurun_out/r2_conv_table9.json --torch-profile gpurun_out/r2_torch_profile9.txt > gpurun_out/r2_bench9.log 2>&1
timeout 300 python bench.py --dtype tf32 --steps 5 > gpurun_out/r2_bench9_tf32.log 2>&1
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k9.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384,8:384:384,1:416:384 > gpurun_out/r2_roles9_l1.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --level 2 --shapes 27:384:384 > gpurun_out/r2_roles9_l2.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --head 3 --shapes 1:256:1024 > gpurun_out/r2_roles9_head.log 2>&1
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range"
$CMD > gpurun_out/r2_plain9.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches9.csv $CMD > gpurun_out/r2_ncu9_a.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 29 -c 15 -o gpurun_out/r2_prof9_tc $CMD > gpurun_out/r2_ncu9_b.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:k_hash_insert|k_assign_rows|k_inverse_accumulate|k_kernel_map_k3_blocks|k_kernel_map_k3$|k_block_rows|k_row_masks|k_mask_keys_rows|k_tile_masks_rows|k_spconv_stem|k_cluster_|k_icp_eval|k_icp_persistent|k_color_|k_select_|k_gather_crops|k_global_pool|k_kabsch|k_translation|k_sanity' -c 40 -o gpurun_out/r2_prof9_misc $CMD > gpurun_out/r2_ncu9_c.log 2>&1
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spconv_tc --csv --log-file gpurun_out/r2_tc_dram9.csv $CMD > gpurun_out/r2_ncu9_d.log 2>&1
tail -3 gpurun_out/r2_pytest9.log; tail -c 300 gpurun_out/r2_bench9.log; ls -la gpurun_out | tail -14
