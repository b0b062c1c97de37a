# integration launch order: the tile list in n parts (device-resident step); 1 = every ds tile from DRAM once per frequency group
rm -f /tmp/ab_ref_tb.npy
for n in 1 0 4 8 12 16 24 32 48 1; do
  RB_RT_PARTS=$n timeout 120 python tools/ab_quick.py parts$n f64 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('parts', $n, 'step', round(d['step_ms'], 4), 'rt', round(d['rt_ms'], 4), 'geo', round(d['geometry_ms'], 4), 'dTb', d.get('max_abs_dTb_K'))"
done
