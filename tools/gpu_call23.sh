mkdir -p gpurun_out
python tools/profile_step.py 4096 2 > gpurun_out/r2q_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sls_mma -c 1 -o gpurun_out/prof_r2q python tools/profile_step.py 4096 2 > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
