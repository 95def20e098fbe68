# C5 and C3 at 8 GPUs (torchrun, one rank per GPU, replicas + one all-gather of counters).  usage: gpurun --gpus 8 --timeout 900 -- 'bash profiles/scripts/run_c5_8gpu.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for w in c5 c3; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $w --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${w}_N${N}.json 2> gpurun_out/${TAG}_bench_${w}_N${N}.err
  tail -c 300 gpurun_out/${TAG}_bench_${w}_N${N}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_${w}_N${N}.json')); print('$w', d['n_gpus'], d['value'], d['unit'], 'e2e', d['e2e']['value'], d['clocks'])
except Exception as e: print('$w failed', e)
PY
done
