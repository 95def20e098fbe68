cd /root/repo
for c in 2 3 4; do for p in 0 50 75 90; do echo "CTAS=$c STATIC=$p"; YGZ_FAST_V2_CTAS=$c YGZ_FAST_STATIC_PCT=$p python profiles/scripts/stage_time.py; done; done
