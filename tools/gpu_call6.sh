mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sls_vs_reference or sls_ragged or closed_loop_vs_oracle or other_horizons or fast_sls_and_qp_class" > gpurun_out/r2_call6_pytest.log 2>&1
tail -5 gpurun_out/r2_call6_pytest.log
{ echo "== latency (sls mma, qp mma)"; python tools/latency_probe.py 1 296 4096; echo "== sls slot kernel"; python tools/latency_probe.py 1 4096 sls_kernel=1; echo "== sls mma warps 4"; python tools/latency_probe.py 4096 sls_mma_warps=4; python tools/latency_probe.py 4096 sls_mma_warps=5;  python tools/latency_probe.py 4096 sls_mma_warps=6;} > gpurun_out/r2_call6_latency.log 2>&1
cat gpurun_out/r2_call6_latency.log
