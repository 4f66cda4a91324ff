mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp_vs_oracle or infeasible or closed_loop_vs_oracle" > gpurun_out/r2_call2_pytest.log 2>&1
tail -5 gpurun_out/r2_call2_pytest.log
{ echo "== latency mma"; python tools/latency_probe.py 1 296 592 4096; echo "== latency slot/block (round 1)"; python tools/latency_probe.py 1 4096 qp_kernel=1; python tools/latency_probe.py 1 296 qp_kernel=2; } > gpurun_out/r2_call2_latency.log 2>&1
cat gpurun_out/r2_call2_latency.log
{ echo "== loop mma"; python tools/iter_stats.py 4096; echo "== loop round-1 kernels"; python tools/iter_stats.py 4096 qp_kernel=1; echo "== loop mma 8192"; python tools/iter_stats.py 8192; } > gpurun_out/r2_call2_loop.log 2>&1
grep -E "TOTAL|==|step  5|step 20" gpurun_out/r2_call2_loop.log
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_run.py block > gpurun_out/r2_call2_racecheck.log 2>&1; tail -8 gpurun_out/r2_call2_racecheck.log
