# Planet.run with host buffers at N=1: copy-out chunks x parts of the integration launch
for cfg in "6 12" "8 12" "12 12" "16 12" "6 6" "6 24" "12 24" "4 12" "6 12"; do
  set -- $cfg
  echo -n "chunks $1 parts $2: "
  RB_RT_CHUNKS=$1 RB_RT_PARTS=$2 timeout 200 python tools/e2e_fill_ab.py 2>&1 | tail -1 | cut -c25-80
done
