mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3g_smoke.log 2>&1; tail -1 gpurun_out/r3g_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; tail -2 gpurun_out/r3g_pytest.log
