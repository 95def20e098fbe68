set -x
cd /root/repo
python profiles/scripts/stage_time.py > gpurun_out/r2a_stage_time.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells" -s 2 -c 1 -o gpurun_out/r2a_fast python profiles/scripts/prof1.py 256 > gpurun_out/r2a_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pyr_resize" -s 14 -c 1 -o gpurun_out/r2a_pyr python profiles/scripts/prof1.py 256 >> gpurun_out/r2a_full.log 2>&1
tail -3 gpurun_out/r2a_full.log
cat gpurun_out/r2a_stage_time.txt
