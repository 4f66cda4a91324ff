mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -5 gpurun_out/r2k_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(json.dumps(d['single_solve_latency'], indent=1))"
