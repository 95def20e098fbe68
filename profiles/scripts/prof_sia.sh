# ncu --set full of k_sia_align in both forms (CTA per problem at P = 592 forced by the knob, warp per problem) + source pages.
# usage: gpurun --timeout 900 -- 'bash profiles/scripts/prof_sia.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
YGZ_SIA_WARP_FROM=100000 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_sia_align -o gpurun_out/${TAG}_sia_cta python profiles/scripts/prof_sia.py 592 > gpurun_out/${TAG}_sia.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_sia_align -o gpurun_out/${TAG}_sia_warp python profiles/scripts/prof_sia.py 592 >> gpurun_out/${TAG}_sia.log 2>&1
tail -3 gpurun_out/${TAG}_sia.log
for f in cta warp; do
  ncu -i gpurun_out/${TAG}_sia_$f.ncu-rep --page source --csv > gpurun_out/${TAG}_sia_${f}_source.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_sia_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_sia_${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_sia_*
