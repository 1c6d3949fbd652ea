mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/r2_pytest3_full.log 2>&1
grep -E "passed|failed|error|Error|differ|assert" gpurun_out/r2_pytest3_full.log | tail -30 > gpurun_out/r2_pytest3.log
timeout 600 python bench.py --stages --torch-profile gpurun_out/r2_torch_profile3.txt --cprofile gpurun_out/r2_cprofile3.txt --conv-table gpurun_out/r2_conv_table3.json > gpurun_out/r2_bench3.log 2>&1
CMD="python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r2_plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches3.csv $CMD > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 42 -c 7 -o gpurun_out/r2_prof_tc $CMD > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_kernel_map_k3_blocks|k_hash_insert|k_spconv_stem|k_cluster_inter|k_icp_persistent|k_block_rows|k_tile_masks_rows|k_mask_keys_rows|k_assign_rows|k_inverse_accumulate' -c 16 -o gpurun_out/r2_prof_misc $CMD > gpurun_out/r2_ncu_c.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_spconv_tc -s 120 -c 120 --csv --log-file gpurun_out/r2_tc_dram3.csv $CMD > gpurun_out/r2_ncu_d.log 2>&1
tail -3 gpurun_out/r2_pytest3.log; tail -c 300 gpurun_out/r2_bench3.log; ls -la gpurun_out | tail -12
