export PM_JIT_W=9 PM_JIT_CTAS=3 PM_JIT_WARPS=8 PM_APX_FAMILY=A
ncu --set full --clock-control none --import-source on -k regex:k_scan_apx --launch-skip 3 -c 1 -o gpurun_out/r02_apx_jit3 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log | cut -c1-300
