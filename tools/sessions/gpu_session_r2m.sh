RB_RT_COMPACT=0 timeout 200 python tools/e2e_ab.py 2>&1 | tail -2
timeout 200 python tools/e2e_ab.py 2>&1 | tail -2
RB_RT_COMPACT=0 timeout 200 python tools/e2e_ab.py 2>&1 | tail -2
timeout 200 python tools/e2e_ab.py 2>&1 | tail -2
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) 2>&1 | tail -4
