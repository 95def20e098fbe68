# Final artefacts of the round in ONE call: GPU tests, C2 bench (own arm + reference arm), ncu launch list + full captures + summary json,
# compute-sanitizer racecheck of smoke(), C5 / C3 lines.  usage: gpurun --timeout 2400 -- 'bash profiles/scripts/prof_r2_final.sh TAG'
TAG=${1:-r2f}
cd /root/repo
bash profiles/scripts/prof_r2_checkpoint.sh $TAG
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv -k regex:k_fast_cells_v2 > gpurun_out/${TAG}_fast_source.csv 2>/dev/null
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${TAG}_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/${TAG}_sanitizer_memcheck.log
[ -n "$SKIP_C3C5" ] || bash profiles/scripts/run_c3c5.sh $TAG

