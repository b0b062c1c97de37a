# sky fill: context stream / side stream / highest-priority side stream, Planet.run on the same box
for m in 0 1 2 0 1 2; do
  RB_FILL_STREAM=$m timeout 200 python tools/e2e_fill_ab.py 2>&1 | tail -1
done
RB_FILL_STREAM=1 RB_TRACE=1 timeout 200 python tools/e2e_ab.py 2>&1 | grep rb_trace | tail -14
