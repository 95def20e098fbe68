# One short call: the GPU test tier; only if it is green, the C2 bench line and the ncu passes that feed profiles/kernel_metrics.json.
# usage: gpurun --timeout 600 -- 'bash profiles/scripts/verify_and_capture.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1
rc=$?; tail -3 gpurun_out/${TAG}_gputests.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED rc=$rc"; exit 1; fi
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench_N1.err; tail -c 200 gpurun_out/${TAG}_bench_N1.err
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_ncu_B256.csv python profiles/scripts/prof1.py 256 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fast_cells_v2|k_orient_describe|k_octree|k_blur_levels|k_bow_match_batch|k_vocab" -o gpurun_out/${TAG}_full python profiles/scripts/prof1.py 256 > gpurun_out/${TAG}_full.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_pyr_resize" -o gpurun_out/${TAG}_full_pyr python profiles/scripts/prof1.py 256 >> gpurun_out/${TAG}_full.log 2>&1
python profiles/scripts/ncu_summary.py 256 gpurun_out/${TAG}_kernel_metrics.json gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_pyr.ncu-rep > /dev/null 2>gpurun_out/${TAG}_summary.err
for r in ${TAG}_full ${TAG}_full_pyr; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_N1.json'))
print({k:d.get(k) for k in ('value','value_single_stream','ms_per_step')}, d['e2e']['value'], d.get('stages_ms_per_step'))
PY
