mkdir -p gpurun_out
for g in 1 2 4; do python bench.py --no-cpu-baseline --groups $g 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('groups', '$g', 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), 'ms/step', round(d['ms_per_step'],3))"; done
for b in 2368 4736 8192; do python bench.py --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch', '$b', 'value', round(d['value']), 'e2e', round(d['e2e']['value'] or 0), 'ms/step', round(d['ms_per_step'],3), d['config']['kernel_ms_avg_over_region'])"; done
