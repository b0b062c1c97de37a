mkdir -p gpurun_out
rm -f /tmp/ab_ref_tb.npy
for srt in "" 1 65536 16384 8192; do
  RB_AB_SORT=$srt timeout 120 python tools/ab_quick.py sort_$srt f64 8 2>&1 | tail -1 | cut -c1-330
done
