mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp_vs_oracle or infeasible or closed_loop_vs_oracle" > gpurun_out/r2_call4_pytest.log 2>&1
tail -3 gpurun_out/r2_call4_pytest.log
{ python tools/qp_phase_profile.py 1; python tools/qp_phase_profile.py 4096; } > gpurun_out/r2_call4_phases.log 2>&1
cat gpurun_out/r2_call4_phases.log
{ echo "== latency mma"; python tools/latency_probe.py 1 296 4096; } > gpurun_out/r2_call4_latency.log 2>&1
cat gpurun_out/r2_call4_latency.log
{ echo "== loop mma"; python tools/iter_stats.py 4096; } > gpurun_out/r2_call4_loop.log 2>&1
grep -E "TOTAL|==|step  5|step 20" gpurun_out/r2_call4_loop.log
