# One call: GPU parity tests, bench lines (own arm + reference arm), ncu launch list and full captures of the step kernels,
# profiles/kernel_metrics.json keyed by the csrc sha.  usage: gpurun --timeout 1500 -- 'bash profiles/scripts/prof_r2_checkpoint.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; tail -4 gpurun_out/${TAG}_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench_N1.err; tail -c 300 gpurun_out/${TAG}_bench_N1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_ncu_B256.csv python profiles/scripts/prof1.py 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fast_cells_v2|k_orient_describe|k_octree|k_blur_levels|k_bow_match_batch|k_vocab" -o gpurun_out/${TAG}_full python profiles/scripts/prof1.py 256 > gpurun_out/${TAG}_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_pyr_resize" -o gpurun_out/${TAG}_full_pyr python profiles/scripts/prof1.py 256 >> gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
python profiles/scripts/ncu_summary.py 256 gpurun_out/${TAG}_kernel_metrics.json gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_pyr.ncu-rep > /dev/null 2>gpurun_out/${TAG}_summary.err
for r in ${TAG}_full ${TAG}_full_pyr; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_N1.json'))
print({k:d.get(k) for k in ('value','value_single_stream','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('stages_ms_per_step'))
print(d.get('cpu_baseline'), d.get('cpu_baseline_opencv'))
print(d['roofline'], d['clocks'])
PY
cat gpurun_out/${TAG}_bench_reference_arm.json | head -c 600
