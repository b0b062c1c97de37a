( timeout 900 python -m pytest tests/test_gpu_alpha.py tests/test_gpu_round2.py -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -4
for cfg in "RB_ALPHA_NO_ORDER=1" "RB_X=1" "RB_ALPHA_NO_ORDER=1" "RB_X=1"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/alpha_c5_probe.py 2>&1 | tail -1 | python -c "
import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(round(d['ms'],3),'ms', round(d['fp64_tflops_algorithmic'],2),'TF', d['max_rel_err_vs_oracle'])"
done
