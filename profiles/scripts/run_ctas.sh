cd /root/repo
mkdir -p gpurun_out
for c in 4 3 2; do
  YGZ_FAST_V2_CTAS=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ctas_$c.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/ctas_$c.json')); print('ctas', $c, 'value', round(d['value']), 'single', round(d['value_single_stream']), 'e2e', round(d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms_per_step'].items()})
PY
done
