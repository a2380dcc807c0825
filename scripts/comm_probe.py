"""Debug probe (torchrun): time of the source-row exchange alone -- uneven all_gather (grouped broadcasts)
vs. even all_gather_into_tensor -- for an [N, F] fp32 matrix split across ranks."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
N, F = 1939743, 128
full = torch.empty(N, F, device='cuda')
per = (N + world - 1) // world
bounds = [min(N, i * per + (37 * i if i and i < world else 0)) for i in range(world)] + [N]   # uneven blocks
own = torch.randn(bounds[rank + 1] - bounds[rank], F, device='cuda')
views = [full[bounds[p]:bounds[p + 1]] for p in range(world)]
even_own = torch.randn(per, F, device='cuda')
even_full = torch.empty(per * world, F, device='cuda')


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t1 = timeit(lambda: dist.all_gather(views, own))
t2 = timeit(lambda: dist.all_gather_into_tensor(even_full, even_own))
if rank == 0:
    mb = N * F * 4 / 1e6
    print('world %d: uneven all_gather %.3f ms, even all_gather_into_tensor %.3f ms (matrix %.0f MB; per-rank receive %.0f MB)'
          % (world, t1, t2, mb, mb * (world - 1) / world))
dist.destroy_process_group()
