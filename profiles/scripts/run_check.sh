# GPU tests + C2 line (+ tuning-knob variants) + C3 line.  usage: gpurun --timeout 1500 -- 'bash profiles/scripts/run_check.sh TAG'
TAG=${1:-chk}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; tail -5 gpurun_out/${TAG}_gputests.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench_N1.err; tail -c 300 gpurun_out/${TAG}_bench_N1.err
YGZ_BLUR_BESIDE_OCTREE=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_N1_blurfork.json 2> gpurun_out/${TAG}_bench_N1_blurfork.err
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c3_N1.json 2> gpurun_out/${TAG}_bench_c3.err; tail -c 300 gpurun_out/${TAG}_bench_c3.err
python - <<PY
import json
for f in ('bench_N1','bench_N1_blurfork'):
    try:
        d=json.load(open('gpurun_out/${TAG}_%s.json'%f)); print(f, d['value'], d.get('value_single_stream'), d['e2e']['value'], d.get('stages_ms_per_step'))
    except Exception as e: print(f,'failed',e)
try:
    d=json.load(open('gpurun_out/${TAG}_bench_c3_N1.json')); print('c3', d['value'], d['e2e']['value'], d['e2e_device_pyramids'])
except Exception as e: print('c3 failed', e)
PY
