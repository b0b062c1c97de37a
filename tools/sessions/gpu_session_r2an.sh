N=${RB_N:-2}
for cfg in "12 6" "6 6" "4 6" "3 6" "2 6" "6 4" "4 4" "3 3"; do
  set -- $cfg
  echo "parts $1 chunks $2:"
  RB_RT_PARTS=$1 RB_RT_CHUNKS=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/e2e_timeline_dist.py 2>&1 | grep -E "^rank" | sort | sed -e 's/prefetch.*rt_batch/rt_batch/' | cut -c1-150
done
