set -x
cd /root/repo
timeout 600 python -m pytest tests/test_extractor_gpu.py tests/test_golden_gpu.py tests/test_dso_gpu.py -x -q 2>&1 | tail -15
python profiles/scripts/stage_time.py
YGZ_FAST_V1=1 python profiles/scripts/stage_time.py
YGZ_FAST_V2_CTAS=3 python profiles/scripts/stage_time.py
