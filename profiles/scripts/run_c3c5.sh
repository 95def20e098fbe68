# C5 and C3 bench lines (own arm + reference arm).  usage: gpurun --timeout 1500 -- 'bash profiles/scripts/run_c3c5.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
for w in c5 c3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${w}_N1.json 2> gpurun_out/${TAG}_bench_${w}.err; tail -c 400 gpurun_out/${TAG}_bench_${w}.err
  timeout 300 python bench.py --workload $w --impl reference > gpurun_out/${TAG}_bench_${w}_reference_arm.json 2>> gpurun_out/${TAG}_bench_${w}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_${w}_N1.json'))
    print('$w', {k:d.get(k) for k in ('value','unit','ms_per_step','gpu_launches','single_problem_latency_ms','gn_iterations_per_s')}, 'e2e', d['e2e']['value'])
    print(d.get('stages_ms_per_step')); print(d.get('cpu_baseline')); print(d['clocks'])
except Exception as e: print('$w failed', e)
PY
done
