mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp or closed_loop_vs_oracle or seeded or drop_in or population or failure" > gpurun_out/r2t_pytest.log 2>&1; tail -2 gpurun_out/r2t_pytest.log
{ timeout 300 python tools/latency_probe.py 4096; } 2>&1 | tail -2
{ timeout 300 python tools/qp_phase_profile.py 4096 qp_kernel=5; } > gpurun_out/r2t_phases.log 2>&1; head -4 gpurun_out/r2t_phases.log
