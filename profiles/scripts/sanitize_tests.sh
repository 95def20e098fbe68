# compute-sanitizer memcheck over the GPU tests of the entry points added late in round 2 (batched DSO selection, SparseImgAlign over
# extractor-resident pyramids, Fuse candidate search, Sim3 family, batched projection searches).
# usage: gpurun --timeout 1500 -- 'bash profiles/scripts/sanitize_tests.sh TAG'
TAG=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
timeout 1300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -x -q -m gpu \
  "tests/test_dso_gpu.py::test_dso_batch_equals_per_frame" \
  "tests/test_sia_gpu.py::test_sia_run_orb_reads_the_extractor_pyramids_in_place" \
  tests/test_fuse_gpu.py tests/test_sim3_gpu.py tests/test_projection_batch_gpu.py tests/test_stereo_gpu.py \
  > gpurun_out/${TAG}_sanitizer_memcheck_tests.log 2>&1
echo "memcheck(tests) rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck_tests.log
