# SparseImgAlign parity tests + C3 line (no CPU arm).  usage: gpurun --timeout 900 -- 'bash profiles/scripts/run_sia_check.sh TAG'
TAG=${1:-chk}
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sia_gpu.py -q 2>&1 | tail -6
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_c3.err; tail -c 300 gpurun_out/${TAG}_c3.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_c3.json')); print(d['value'], d['single_problem_latency_ms'], d['e2e']['value'], d['e2e_device_pyramids']['value'])"
