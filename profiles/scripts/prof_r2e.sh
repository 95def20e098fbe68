cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:"k_pyr_resize_v2" -s 14 -c 2 -o gpurun_out/r2e_pyr python profiles/scripts/prof1.py 256 > gpurun_out/r2e_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pyr" -s 14 -c 7 --csv --log-file gpurun_out/r2e_pyr_launches.csv python profiles/scripts/prof1.py 256 > /dev/null 2>&1
grep -o '"k_pyr[^"]*".*' gpurun_out/r2e_pyr_launches.csv | awk -F'","' '{print $4, $5, $NF}'
