mkdir -p gpurun_out
for cfg in "RB_ALPHA_BLOCKS_PER_LAYER=0" "RB_ALPHA_BLOCKS_PER_LAYER=4" "RB_ALPHA_BLOCKS_PER_LAYER=2" "RB_ALPHA_BLOCKS_PER_LAYER=1" "RB_ALPHA_FPT=4 RB_ALPHA_BLOCKS_PER_LAYER=1" "RB_ALPHA_FPT=4 RB_ALPHA_BLOCKS_PER_LAYER=2"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/alpha_c5_probe.py 2>&1 | tail -1 | python -c "
import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(round(d['ms'],3),'ms', round(d['fp64_tflops_algorithmic'],2),'TF', d['max_rel_err_vs_oracle'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"alpha_lines" -s 3 -c 1 -f -o gpurun_out/prof_r2g_alpha python tools/alpha_c5_probe.py > gpurun_out/r2g_ncu.log 2>&1
ls -la gpurun_out/prof_r2g_alpha.ncu-rep
