mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3f_smoke.log 2>&1; tail -1 gpurun_out/r3f_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; tail -2 gpurun_out/r3f_pytest.log
python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; tail -c 300 gpurun_out/r3f_bench.json
python bench.py --impl reference > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_bench_reference.err; tail -c 200 gpurun_out/r3f_bench_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3f_ncu.log 2>&1
tail -2 gpurun_out/r3f_ncu.log; wc -l gpurun_out/r3f_launches.csv
