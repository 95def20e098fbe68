set -x
cd /root/repo
# launch list of one bench-shaped step (B=256): every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s 22 -c 12 --csv --log-file gpurun_out/r1_launches.csv python profiles/scripts/prof1.py 256 > /dev/null 2>&1
# full capture of every kernel of the step
ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells|k_orient_describe|k_octree|k_blur_levels|k_bow_match_batch" -s 8 -c 5 -o gpurun_out/r1_full python profiles/scripts/prof1.py 256 > gpurun_out/r1_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pyr_resize" -s 14 -c 7 -o gpurun_out/r1_full_pyr python profiles/scripts/prof1.py 256 >> gpurun_out/r1_full.log 2>&1
tail -3 gpurun_out/r1_full.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_N1.json 2> gpurun_out/r1_bench_N1.err
tail -c 400 gpurun_out/r1_bench_N1.json
