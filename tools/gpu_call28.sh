mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; tail -2 gpurun_out/r3e_pytest.log
python bench.py > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; tail -c 600 gpurun_out/r3e_bench.json
python bench.py --impl reference > gpurun_out/r3e_bench_reference.json 2> gpurun_out/r3e_bench_reference.err; tail -c 300 gpurun_out/r3e_bench_reference.json
python tools/profile_step.py 4096 2 > gpurun_out/r3e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"qp_tm_kernel|sls_mma_kernel" -c 3 -o gpurun_out/prof_r3e python tools/profile_step.py 4096 2 > gpurun_out/r3e_ncu.log 2>&1
tail -2 gpurun_out/r3e_ncu.log
