mkdir -p gpurun_out
python tools/profile_step.py 4096 2 > gpurun_out/r2i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_tm -s 1 -c 2 -o gpurun_out/prof_r2i python tools/profile_step.py 4096 2 > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_plain.log; tail -3 gpurun_out/r2i_ncu.log
