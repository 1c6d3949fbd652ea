mkdir -p gpurun_out
P=$PWD/markerless-robot-camera-calibration_b200
V=$P/lib_variants
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest11_full.log 2>&1
tail -5 gpurun_out/r2_pytest11_full.log > gpurun_out/r2_pytest11.log
PR="timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep"
$PR --lib-b $V/libb2me_trim.so --shapes 27:384:384,8:384:384,1:416:384,1:384:256 > gpurun_out/r2_probe11_l1.log 2>&1
$PR --lib-b $V/libb2me_trim.so --head 3 --shapes 1:256:1024 > gpurun_out/r2_probe11_head.log 2>&1
$PR --lib-b $V/libb2me_trim.so --level 2 --shapes 27:384:384,27:32:32,8:384:384 > gpurun_out/r2_probe11_l2.log 2>&1
$PR --lib-b $V/libb2me_k8split.so --shapes 8:384:384 > gpurun_out/r2_probe11_k8split.log 2>&1
$PR --lib-b $V/libb2me_k8split.so --level 2 --shapes 8:384:384 > gpurun_out/r2_probe11_k8split_l2.log 2>&1
timeout 600 python bench.py --stages --crop both --conv-table gpurun_out/r2_conv_table11.json --torch-profile gpurun_out/r2_torch_profile11.txt > gpurun_out/r2_bench11.log 2>&1
timeout 600 python bench.py --strong-frames 96 --steps 2 --warmup 1 > gpurun_out/r2_strong11.log 2>&1
B2ME_LIB_PATH=$P/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 8:384:384,1:416:384 > gpurun_out/r2_roles11_l1.log 2>&1
du -sh gpurun_out; tail -3 gpurun_out/r2_pytest11.log; grep -h "median" gpurun_out/r2_probe11_*.log | grep -E "default" | cut -c1-90 ; tail -c 300 gpurun_out/r2_bench11.log
