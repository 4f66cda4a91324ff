# round-2 baseline of the committed state: gpu tests, bench line, launch list, full capture of the QP kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
python tools/profile_step.py 4096 2 > gpurun_out/r2a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python tools/profile_step.py 4096 2 > gpurun_out/r2a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qp_mma -s 12 -c 2 -o gpurun_out/prof_r2a python tools/profile_step.py 4096 2 > gpurun_out/r2a_ncu2.log 2>&1
tail -3 gpurun_out/r2a_plain.log
