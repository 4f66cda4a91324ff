mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
python bench.py --impl reference > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; tail -c 700 gpurun_out/r2j_bench_reference.json
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 1500 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
python tools/profile_step.py 4096 2 > gpurun_out/r2j_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches.csv python tools/profile_step.py 4096 2 > gpurun_out/r2j_ncu1.log 2>&1
