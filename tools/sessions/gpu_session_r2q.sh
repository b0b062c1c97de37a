( RB_RT_TILES=1 timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -5
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -5
