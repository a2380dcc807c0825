"""torch.profiler breakdown of the neighbour-sampled train step (tuning aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from re_gnn_b200 import Graph, mag, synth
from re_gnn_b200.sampling import NeighborSampler
d = synth.hetero_graph('mag'); dev = 'cuda:0'
n, net, sizes = d['num_nodes'], d['num_etype'], d['type_sizes']; nnt = len(sizes)
g = Graph(d['src'][:-n], d['dst'][:-n], n).to(dev); g.csr()
edge_type0 = (torch.as_tensor(d['etype'][:-n]) - 1).to(dev); node_type = torch.as_tensor(d['ntype']).to(dev)
offs = np.concatenate([[0], np.cumsum(sizes)]); local_idx = torch.as_tensor(np.arange(n) - offs[d['ntype']]).to(dev)
x_dict = {k: torch.randn(sizes[k], 128, device=dev) for k in range(nnt)}
labels = torch.randint(0, 349, (n,), device=dev)
model = mag.REGNN(128, 512, 349, 1, 2, 100.0, 0.5, {k: 128 for k in range(nnt)}, net, residual=True, no_re=False).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
sampler = NeighborSampler(g, [25, 20], seed=1)
def step(i):
    seeds = torch.randperm(600000, device=dev)[:512]
    return mag.train_step(model, opt, sampler, seeds, labels[seeds], x_dict, edge_type0, node_type, local_idx, 0, i)
for i in range(5): step(i)
torch.cuda.synchronize()
# phase timing
t0 = time.perf_counter()
for i in range(10):
    seeds = torch.randperm(600000, device=dev)[:512]
    n_id, blocks = sampler.sample(seeds, 0, i)
torch.cuda.synchronize(); print('sample only: %.2f ms/step' % ((time.perf_counter() - t0) * 100))
t0 = time.perf_counter()
for i in range(10): step(i)
torch.cuda.synchronize(); print('full step: %.2f ms/step' % ((time.perf_counter() - t0) * 100))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=14, max_name_column_width=60))
