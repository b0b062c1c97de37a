N=${1:-8}
for cw in 0 0.15 0.3 0.6; do
  RB_TRACE_SKIP_COPIES=1 RB_ROW_COPY_WEIGHT=$cw timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/e2e_trace_dist.py 2>&1 | tr "\r" "\n" | grep -E "^rank (0|3|7) " | sort
done
