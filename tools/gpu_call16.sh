mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s -k "population or drop_in" > gpurun_out/r2l_pytest.log 2>&1; grep -E "^population|passed|failed|Error|assert" gpurun_out/r2l_pytest.log | head -30
