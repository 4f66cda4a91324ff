mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sls or closed_loop_vs_oracle or seeded" > gpurun_out/r2r_pytest.log 2>&1; tail -2 gpurun_out/r2r_pytest.log
{ timeout 300 python tools/latency_probe.py 4096 qp_tm_ctas=2; } 2>&1 | tail -2
{ timeout 300 python tools/latency_probe.py 4096 qp_tm_ctas=1; } 2>&1 | tail -2
{ timeout 300 python tools/latency_probe.py 2368 qp_tm_ctas=2; } 2>&1 | tail -1
{ timeout 300 python tools/latency_probe.py 1184 qp_tm_ctas=2; } 2>&1 | tail -1
{ timeout 300 python tools/latency_probe.py 592 qp_tm_ctas=1 qp_kernel=5; } 2>&1 | tail -1
for g in 1 2; do python bench.py --no-cpu-baseline --groups $g 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('groups', '$g', 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), 'ms/step', round(d['ms_per_step'],3), d['config'].get('kernel_ms_avg_over_region'))"; done
