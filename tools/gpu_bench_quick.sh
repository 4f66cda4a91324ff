python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), d['config']['kernel_ms_avg_over_region'], 'e2e_invalid', d['e2e'].get('invalid'))"
