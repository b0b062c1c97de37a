( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -6
rm -f /tmp/ab_ref_tb.npy
RB_RT_PARTS=0 timeout 120 python tools/ab_quick.py tile_major f64 8 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/ab_quick.py fg_major f64 8 2>&1 | tail -1 | cut -c1-400
for p in 0 6 12 24 48; do echo "parts $p"; RB_RT_PARTS=$p timeout 200 python tools/e2e_ab.py 2>&1 | tail -2 | head -1; done
