for w in 13 12 10 7; do echo "== warps $w"; QS_WARPS_PER_CTA=$w python scripts/perf_probe.py cfg2 --steps 500 2>&1 | grep -v '"variant": "f3"' ; done
