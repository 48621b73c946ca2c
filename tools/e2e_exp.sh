for pb in 2048 613 409 307; do
  SFM_E2E_PAIR_BATCH=$pb timeout 300 python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$pb', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'])"
done
