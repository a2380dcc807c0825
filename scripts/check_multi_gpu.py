"""Multi-rank GPU parity check (run under torchrun, one rank per GPU): the three distributed aggregation schemes of
re_gnn_b200.partition -- row blocks + all-gather ('rows'), column slabs + NCCL all-to-all ('cols'), column slabs over
peer memory inside our kernels ('peer') -- against the single-device functional.propagate on every rank.
Outputs and dX must be bit-identical (row sums run in slot order everywhere); the relation gradient, summed over
ranks, within 1e-5 relative."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from re_gnn_b200 import Graph, functional as RF, partition, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    feat = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    d = synth.hetero_graph('mag', seed=3, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 0.05)
    n, r = d['num_nodes'], d['num_relations']
    g = Graph(d['src'], d['dst'], n).to(dev)
    etv = g.etype_views(torch.as_tensor(d['etype']).to(dev), r)
    gen = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(n, feat, device=dev, generator=gen)
    gout = torch.randn(n, feat, device=dev, generator=gen)
    theta0 = (torch.rand(r, 1, device=dev, generator=gen) + 0.5) / 100

    def single():
        xs, th = x.clone().requires_grad_(True), theta0.clone().requires_grad_(True)
        out = RF.propagate(g, etv, xs, th, 100.0, RF.weighted_degree_norm(g, etv, th, 100.0, -0.5))
        out.backward(gout)
        return out.detach(), xs.grad, th.grad
    ref = single()
    ok_all = True
    for mode in ['rows', 'cols', 'peer']:
        bounds = partition.row_blocks(g.csr()['indptr'], world, balance='rows')
        rb, re = bounds[rank], bounds[rank + 1]
        xch = partition.SlabExchange(feat, bounds, rank, dev) if mode == 'peer' else None
        for rep in range(2):   # twice: the exchange buffers are reused
            xo, th = x[rb:re].clone().requires_grad_(True), theta0.clone().requires_grad_(True)
            nrm = RF.weighted_degree_norm(g, etv, th, 100.0, -0.5)
            if mode == 'rows':
                out = partition.partitioned_propagate(g, etv, xo, th, 100.0, nrm, bounds, rank)
            else:
                out = partition.feature_sliced_propagate(g, etv, xo, th, 100.0, nrm, bounds, rank, exchange=xch,
                                                         alias=(rep == 1))
            out.backward(gout[rb:re])
            partition.allreduce_relation_grads([th], exchange=xch)
            e_out = bool(torch.equal(out.detach(), ref[0][rb:re]))
            e_dx = bool(torch.equal(xo.grad, ref[1][rb:re]))
            rel = float((th.grad - ref[2]).abs().max() / ref[2].abs().max())
            ok = e_out and e_dx and rel < 1e-5
            flag = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print('%s rep%d: out_equal=%s dx_equal=%s d_theta_rel=%.2e all_ranks_ok=%d' % (mode, rep, e_out, e_dx, rel,
                                                                                              int(flag.item())), flush=True)
            ok_all = ok_all and bool(flag.item())
    if rank == 0:
        print('MULTI_GPU_CHECK', 'OK' if ok_all else 'FAIL', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == '__main__':
    main()
