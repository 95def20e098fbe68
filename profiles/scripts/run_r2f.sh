cd /root/repo
timeout 900 python -m pytest tests/test_vocab_gpu.py tests/test_matcher_gpu.py -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 600 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:d[k] for k in ('value','value_single_stream','ms_per_step','gpu_launches')}, d['e2e']['value'], d['stages_ms_per_step'])
print(d.get('cpu_baseline'), d.get('cpu_baseline_opencv'))
print(d['roofline']['frac'], d['roofline']['limiter'], d['clocks'])
PY
