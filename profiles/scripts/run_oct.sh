cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_extractor_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -3
for u in 8 4; do echo "YGZ_OCT_UNROLL=$u"; YGZ_OCT_UNROLL=$u python profiles/scripts/stage_time.py; YGZ_OCT_UNROLL=$u python profiles/scripts/stage_time.py; done
