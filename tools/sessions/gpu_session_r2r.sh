# 2 GPUs: the in-kernel all_gather of the layer-sharded absorption against kernel + NCCL all_gather
N=${1:-2}
cat > /tmp/alpha_gather_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
import bench
from radiobear_b200 import _lib
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
ctx = _lib.get_context(local)
world, rank = dist.get_world_size(), dist.get_rank()
for mode in ('nccl', 'kernel', 'nccl', 'kernel'):
    os.environ['RB_BENCH_ALPHA_GATHER'] = mode
    r = bench.alpha_c5(ctx, dev, world=world, rank=rank)
    if rank == 0:
        print(mode, r['gather'], 'ms', round(r['ms'], 4), 'kernel', round(r['kernel_ms'], 4), 'err', r['max_rel_err_vs_oracle'], flush=True)
# every rank holds the same slab?
sym = __import__('radiobear_b200.parallel', fromlist=['x']).symmetric_slab(4096, 4096, dev)
if sym.usable:
    cs = sym.tensor.sum().reshape(1)
    lst = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(lst, cs)
    if rank == 0:
        print('checksums equal on all ranks:', all(bool(x.item() == lst[0].item()) for x in lst))
else:
    print('symmetric slab unusable:', getattr(sym, 'error', None))
dist.barrier(); dist.destroy_process_group()
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 /tmp/alpha_gather_ab.py 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*" | tail -12
