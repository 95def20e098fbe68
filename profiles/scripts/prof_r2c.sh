set -x
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells_v2" -s 2 -c 1 -o gpurun_out/r2c_fast python profiles/scripts/prof1.py 256 > gpurun_out/r2c_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fast" -s 4 -c 4 --csv --log-file gpurun_out/r2c_fast_launches.csv python profiles/scripts/prof1.py 256 > /dev/null 2>&1
cat gpurun_out/r2c_fast_launches.csv | tail -6
