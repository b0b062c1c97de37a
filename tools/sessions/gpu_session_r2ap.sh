# Doppler path: per-layer frequencies in the absorption kernel, slab pair in the integration; then the whole GPU suite
( timeout 600 python -m pytest tests/test_gpu_alpha.py tests/test_gpu_planet.py -m gpu -q --tb=short -k "per_layer or doppler"; echo "pytest exit $?" ) 2>&1 | tail -30
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
