# final ncu captures of round 2 (final build: TMA operand path, K1 table lines, ICP pre-filter); exports CSV on the box
mkdir -p gpurun_out
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range"
$CMD > gpurun_out/r2_plain25.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches25.csv $CMD > gpurun_out/r2_ncu25_a.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 40 -c 9 -o gpurun_out/r2_prof25_tc $CMD > gpurun_out/r2_ncu25_b.log 2>&1
R=gpurun_out/r2_prof25_tc.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page details --csv > gpurun_out/r2_prof25_tc_details.csv 2>/dev/null
  ncu -i $R --page raw --csv > gpurun_out/r2_prof25_tc_raw.csv 2>/dev/null
  rm -f $R
fi
ncu --profile-from-start off --set full --clock-control none -k 'regex:k_hash_insert|k_assign_rows|k_inverse_accumulate|k_kernel_map_k3_blocks|k_kernel_map_k3$|k_block_rows|k_row_masks|k_mask_keys_rows|k_tile_masks_rows|k_color_|k_first_flags|k_stride_kernel_maps|k_quantize' -c 40 -o gpurun_out/r2_prof25_misc $CMD > gpurun_out/r2_ncu25_c.log 2>&1
R=gpurun_out/r2_prof25_misc.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page raw --csv > gpurun_out/r2_prof25_misc_raw.csv 2>/dev/null
  rm -f $R
fi
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spconv_tc --csv --log-file gpurun_out/r2_tc_dram25.csv $CMD > gpurun_out/r2_ncu25_d.log 2>&1
du -sh gpurun_out; ls gpurun_out
