N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/e2e_timeline_dist.py 2>&1 | grep -E "^rank" | sort
RB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 tools/e2e_timeline_dist.py 2>&1 | grep -E "rb_trace" | tail -12
