rm -f /tmp/ab_ref_tb.npy
for v in cur w8 c5 s10 w2; do
  RB_LIB_PATH=radiobear_b200/lib/librb_$v.so timeout 120 python tools/ab_quick.py $v f64 8 2>&1 | tail -1 | cut -c1-250
done
