mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "soft_constrained or compare_driver" > gpurun_out/r2n_pytest.log 2>&1; grep -E "closed-loop cost|passed|failed|Error|assert|status" gpurun_out/r2n_pytest.log | head -30
