mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2o_pytest.log 2>&1; grep -E "first .* steps|passed|failed|Error" gpurun_out/r2o_pytest.log | head; tail -2 gpurun_out/r2o_pytest.log
