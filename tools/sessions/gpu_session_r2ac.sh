N=${1:-8}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/e2e_trace_dist.py 2>&1 | grep -E "^rank|rb_trace" ; }
RB_TRACE=1 run | tail -12
RB_RT_CHUNKS=1 run | sort | head -8
RB_RT_CHUNKS=3 run | sort | head -8
RB_RT_CHUNKS=12 run | sort | head -8
