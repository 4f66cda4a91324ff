mkdir -p gpurun_out
python tools/latency_probe.py 296 > gpurun_out/r2_call5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qp_mma_kernel -s 8 -c 1 -o gpurun_out/prof_qpmma_v3 python tools/latency_probe.py 296 > gpurun_out/r2_call5_ncu.log 2>&1
tail -3 gpurun_out/r2_call5_ncu.log
