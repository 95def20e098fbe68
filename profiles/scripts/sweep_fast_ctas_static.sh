# FAST kernel tuning sweep: resident CTAs per SM x static share of the work queue (stage timing by stage_time.py).  usage: gpurun -- bash profiles/scripts/sweep_fast_ctas_static.sh
cd /root/repo
for c in 2 3 4; do for p in 0 50 75 90; do echo "CTAS=$c STATIC=$p"; YGZ_FAST_V2_CTAS=$c YGZ_FAST_STATIC_PCT=$p python profiles/scripts/stage_time.py; done; done
