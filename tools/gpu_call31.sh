mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qp or closed_loop_vs_oracle or seeded or drop_in or population or failure" > gpurun_out/r2t_pytest.log 2>&1; tail -1 gpurun_out/r2t_pytest.log
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), d['config']['kernel_ms_avg_over_region'])"
{ timeout 300 python tools/qp_phase_profile.py 4096 qp_kernel=5; } > gpurun_out/r2t_phases.log 2>&1; sed -n 1,16p gpurun_out/r2t_phases.log
