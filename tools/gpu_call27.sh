mkdir -p gpurun_out
python tools/profile_step.py 4096 4 > gpurun_out/r2u_plain.log 2>&1 && tail -3 gpurun_out/r2u_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_raw.csv python tools/profile_step.py 4096 4 > gpurun_out/r2u_ncu1.log 2>&1
tail -1 gpurun_out/r2u_ncu1.log
