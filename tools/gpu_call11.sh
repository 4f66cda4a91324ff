mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp_vs_oracle or closed_loop_vs_oracle or infeasible or seeded or other_horizons" > gpurun_out/r2h_pytest.log 2>&1; tail -4 gpurun_out/r2h_pytest.log
{ echo "== tmem (auto)"; timeout 300 python tools/latency_probe.py 1024 4096; echo "== warp"; timeout 300 python tools/latency_probe.py 4096 qp_kernel=4; } > gpurun_out/r2h_latency.log 2>&1; cat gpurun_out/r2h_latency.log
{ timeout 300 python tools/qp_phase_profile.py 4096 qp_kernel=5; } > gpurun_out/r2h_phases.log 2>&1; cat gpurun_out/r2h_phases.log
{ echo "== loop tmem"; timeout 300 python tools/iter_stats.py 4096; } > gpurun_out/r2h_loop.log 2>&1
grep -E "TOTAL|==|step  5|step 20" gpurun_out/r2h_loop.log
