mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3h_smoke.log 2>&1; tail -1 gpurun_out/r3h_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; tail -2 gpurun_out/r3h_pytest.log
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
