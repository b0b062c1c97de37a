# N = 2 on the final code: sharded image / absorption against the single-GPU results, then the bench line
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL" | tail -8
bash tools/sessions/gpu_session_scale.sh 2
