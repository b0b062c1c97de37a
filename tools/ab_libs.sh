# A/B several builds of the library on the GPU box: tools/ab_libs.sh <variant> ... (radiobear_b200/lib/librb_<variant>.so)
for v in "$@"; do
  cp radiobear_b200/lib/librb_$v.so radiobear_b200/lib/libradiobear_b200.so
  timeout 200 python bench.py --steps 6 --warmup 3 > gpurun_out/ab_$v.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/ab_$v.json')); print('$v', round(d['ms_per_step'],3), round(d['kernels_ms']['rt_integrate'],3))"
done
