"""Gradient all-reduce of the data-parallel minibatch path over gloo (world size 2, CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from re_gnn_b200 import mag
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    x = torch.full((4, 3), float(rank + 1))
    lin(x).sum().backward()
    n = mag.allreduce_gradients(list(lin.parameters()), world)
    # rank r contributes gradients proportional to (r+1): the average over ranks 0,1 is 1.5x rank 0's
    want_w = torch.full((2, 3), 4.0 * 1.5)
    ret[rank] = bool(n == 8 and torch.allclose(lin.weight.grad, want_w) and torch.allclose(lin.bias.grad, torch.full((2,), 4.0)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_allreduce_world2():
    port = 31500 + (os.getpid() % 2000)
    ret = mp.get_context('spawn').Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2)), dict(ret)
