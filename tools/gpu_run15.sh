mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest15_full.log 2>&1
tail -5 gpurun_out/r2_pytest15_full.log > gpurun_out/r2_pytest15.log
timeout 600 python bench.py --stages --no-cpu-baseline > gpurun_out/r2_bench15.log 2>&1
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range"
ncu --profile-from-start off --set full --clock-control none -k 'regex:k_hash_insert|k_assign_rows|k_inverse_accumulate|k_kernel_map_k3_blocks|k_kernel_map_k3$|k_block_rows|k_row_masks|k_mask_keys_rows|k_tile_masks_rows|k_spconv_stem|k_cluster_|k_icp_eval|k_color_|k_select_|k_gather_crops|k_global_pool|k_kabsch|k_translation|k_sanity|k_first_flags|k_stride_kernel_maps|k_quantize' -c 40 -o gpurun_out/r2_prof15_misc $CMD > gpurun_out/r2_ncu15_c.log 2>&1
R=gpurun_out/r2_prof15_misc.ncu-rep
if [ -f $R ]; then
  ncu -i $R --page raw --csv > gpurun_out/r2_prof15_misc_raw.csv 2>/dev/null
  rm -f $R
fi
tail -3 gpurun_out/r2_pytest15.log; tail -c 400 gpurun_out/r2_bench15.log
