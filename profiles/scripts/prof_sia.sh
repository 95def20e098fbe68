# ncu --set full of k_sia_align on a C3-shaped launch (592 problems) + source / raw pages.
# usage: gpurun --timeout 900 -- 'bash profiles/scripts/prof_sia.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_sia_align -o gpurun_out/${TAG}_sia python profiles/scripts/prof_sia.py 592 > gpurun_out/${TAG}_sia.log 2>&1
tail -3 gpurun_out/${TAG}_sia.log
ncu -i gpurun_out/${TAG}_sia.ncu-rep --page source --csv > gpurun_out/${TAG}_sia_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_sia.ncu-rep --page raw --csv > gpurun_out/${TAG}_sia_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_sia*
