# ncu only: launch list + full captures of one bench-shaped step (prof1.py brackets it with cudaProfilerStart/Stop) + summary json
# + per-source-line hot spots of the FAST kernel.  usage: gpurun --timeout 900 -- 'bash profiles/scripts/prof_r2_ncu.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_ncu_B256.csv python profiles/scripts/prof1.py 256 > gpurun_out/${TAG}_prof1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_fast_cells_v2|k_orient_describe|k_octree|k_blur_levels|k_bow_match_batch|k_vocab" -o gpurun_out/${TAG}_full python profiles/scripts/prof1.py 256 > gpurun_out/${TAG}_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_pyr_resize" -o gpurun_out/${TAG}_full_pyr python profiles/scripts/prof1.py 256 >> gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
python profiles/scripts/ncu_summary.py 256 gpurun_out/${TAG}_kernel_metrics.json gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_pyr.ncu-rep > /dev/null 2>gpurun_out/${TAG}_summary.err
for r in ${TAG}_full ${TAG}_full_pyr; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv -k regex:k_fast_cells_v2 > gpurun_out/${TAG}_fast_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv -k regex:k_octree > gpurun_out/${TAG}_octree_source.csv 2>/dev/null
cat gpurun_out/${TAG}_launches_ncu_B256.csv | grep -o 'k_[a-z_0-9]*(.*' | awk -F'","' '{print $1, $NF}' | cut -c1-40,200-
