( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -8
rm -f /tmp/ab_ref_tb.npy
for i in 1 2; do timeout 120 python tools/ab_quick.py squared f64 8 2>&1 | tail -1 | cut -c1-360; done
