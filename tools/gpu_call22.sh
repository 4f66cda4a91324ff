mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp_vs_oracle or closed_loop_vs_oracle or sls_vs_reference or sls_ragged or other_horizons or seeded" > gpurun_out/r2p_pytest.log 2>&1; tail -2 gpurun_out/r2p_pytest.log
{ timeout 300 python tools/latency_probe.py 4096; } 2>&1 | tail -3
{ timeout 300 python tools/qp_phase_profile.py 4096 qp_kernel=5; } > gpurun_out/r2p_phases.log 2>&1; cat gpurun_out/r2p_phases.log
timeout 300 python tools/iter_stats.py 4096 | tail -1
