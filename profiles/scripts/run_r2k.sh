# GPU tests + compute-sanitizer pass of smoke() + C5 re-measure.  usage: gpurun --timeout 1800 -- 'bash profiles/scripts/run_r2k.sh TAG'
TAG=${1:-r2k}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; tail -4 gpurun_out/${TAG}_gputests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5_N1.json 2> gpurun_out/${TAG}_bench_c5.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_c5_N1.json')); print('c5', d['value'], 'e2e', d['e2e'])
PY
