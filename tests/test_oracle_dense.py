"""Second, independent pin of the oracle: a dense-adjacency formulation on tiny graphs (float64),
analytic known answers and autograd gradcheck."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import csr_oracle, regnn_oracle as O
from re_gnn_b200 import synth


def tiny(seed=0, n=12, e=40, r=5):
    d = synth.random_multigraph(n, e, r, seed=seed)
    return (torch.as_tensor(d['src']), torch.as_tensor(d['dst']), torch.as_tensor(d['etype']), n, r)


def dense_adj(src, dst, w, n):
    """A[v, u] = sum of w over edges u -> v (multi-edges add up)."""
    a = torch.zeros(n, n, dtype=torch.float64)
    a.index_put_((dst, src), w, accumulate=True)
    return a


def test_regcn_equals_dense_formulation():
    src, dst, et, n, r = tiny(1)
    th = torch.as_tensor(np.random.RandomState(0).uniform(-0.5, 1.5, (r, 1)) / 100.0)
    x = torch.randn(n, 6, dtype=torch.float64)
    w = F.leaky_relu(th * 100.0, 0.01)[et - 1, 0]
    a = dense_adj(src, dst, w, n)
    nrm = a.sum(1).clamp(min=1) ** -0.5
    want = nrm[:, None] * (a @ (nrm[:, None] * x))
    got = O.regraphconv_forward(src, dst, et, n, x, th, 100.0)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_regcn_at_init_is_symmetric_normalised_gcn():
    """edge_weight*alpha == 1 at init => D^-1/2 A^T D^-1/2 X with D the in-degree incl. self loops."""
    src, dst, et, n, r = tiny(2)
    th = torch.full((r, 1), 1.0 / 100.0, dtype=torch.float64)
    x = torch.randn(n, 4, dtype=torch.float64)
    a = dense_adj(src, dst, torch.ones(src.numel(), dtype=torch.float64), n)
    deg = torch.as_tensor(csr_oracle.in_degrees(dst.numpy(), n)).double()
    assert torch.equal(a.sum(1), deg)
    nrm = deg.clamp(min=1) ** -0.5
    want = nrm[:, None] * (a @ (nrm[:, None] * x))
    assert torch.allclose(O.regraphconv_forward(src, dst, et, n, x, th, 100.0), want, rtol=1e-12, atol=1e-12)


def test_regat_equal_relations_is_plain_gat_and_dense_softmax():
    src, dst, et, n, r = tiny(3)
    h, dd = 2, 3
    f = torch.randn(n, h, dd, dtype=torch.float64)
    al, ar = torch.randn(1, h, dd, dtype=torch.float64), torch.randn(1, h, dd, dtype=torch.float64)
    th = torch.full((r, h), 0.37 / 100.0, dtype=torch.float64)     # same embedding for every type
    got, att = O.regat_forward(src, dst, et, n, f.reshape(n, -1), al, ar, th, 100.0, 0.2, return_attention=True)
    plain = O.regat_forward(src, dst, None, n, f.reshape(n, -1), al, ar, th, 100.0, 0.2)
    # softmax is shift-invariant only before the LeakyReLU, so compare against the explicit formula
    el, er = (f * al).sum(-1), (f * ar).sum(-1)
    want = torch.zeros_like(f)
    for v in range(n):
        idx = (dst == v).nonzero().view(-1)
        if idx.numel() == 0:
            continue
        l = F.leaky_relu(el[src[idx]] + er[v] + 0.37, 0.2)
        a = torch.softmax(l, 0)
        want[v] = (a[:, :, None] * f[src[idx]]).sum(0)
        assert torch.allclose(att[idx, :, 0], a, rtol=1e-12, atol=1e-14)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
    assert plain.shape == got.shape


def test_regatv2_row_by_row():
    src, dst, et, n, r = tiny(4)
    h, dd = 2, 4
    fs, fd = torch.randn(n, h, dd, dtype=torch.float64), torch.randn(n, h, dd, dtype=torch.float64)
    at = torch.randn(1, h, dd, dtype=torch.float64)
    th = torch.as_tensor(np.random.RandomState(1).uniform(-0.5, 1.5, (r, h)) / 100.0)
    w = F.leaky_relu(th * 100.0, 0.01)
    src_l, dst_l = src, dst
    e = (F.leaky_relu(fs[src_l] + fd[dst_l], 0.2) * at).sum(-1) + w[et - 1]
    a = O.edge_softmax(e, dst_l, n)
    got = O.segment_sum(fs[src_l] * a[:, :, None], dst_l, n)
    want = torch.zeros_like(fs)
    for v in range(n):
        idx = (dst == v).nonzero().view(-1)
        if idx.numel():
            l = (F.leaky_relu(fs[src[idx]] + fd[v], 0.2) * at).sum(-1) + w[et[idx] - 1]
            want[v] = (torch.softmax(l, 0)[:, :, None] * fs[src[idx]]).sum(0)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_remixhop_dense_powers():
    src, dst, et, n, r = tiny(5)
    th = torch.as_tensor(np.random.RandomState(2).uniform(0.5, 1.5, (r, 1)) / 100.0)
    x = torch.randn(n, 5, dtype=torch.float64)
    ws = {j: torch.randn(3, 5, dtype=torch.float64) for j in (0, 1, 2)}
    w = F.leaky_relu(th * 100.0, 0.01)[et - 1, 0]
    nrm = dense_adj(src, dst, w, n).sum(1).clamp(min=1) ** -0.5
    a1 = dense_adj(src, dst, torch.ones(src.numel(), dtype=torch.float64), n)     # un-weighted propagation
    ahat = nrm[:, None] * a1 * nrm[None, :]
    want = torch.cat([x @ ws[0].t(), (ahat @ x) @ ws[1].t(), (ahat @ ahat @ x) @ ws[2].t()], 1)
    got = O.remixhop_forward(src, dst, et, n, x, th, 100.0, ws, (0, 1, 2))
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_clamp_passes_gradient_at_degree_exactly_one():
    """An isolated node with only a self loop of weight exactly 1 (Appendix A.2)."""
    src, dst, et = torch.tensor([0, 1]), torch.tensor([0, 1]), torch.tensor([1, 1])
    th = torch.tensor([[0.01]], dtype=torch.float64, requires_grad=True)
    x = torch.ones(2, 1, dtype=torch.float64)
    O.regraphconv_forward(src, dst, et, 2, x, th, 100.0).sum().backward()
    # out = w * n^2 = w / max(w,1); at w=1 the clamp passes the gradient: d/dw = 1 - 1 = 0 per node
    assert torch.allclose(th.grad, torch.zeros_like(th.grad), atol=1e-12)


def test_oracle_gradcheck():
    src, dst, et, n, r = tiny(6, n=6, e=14, r=3)
    th = torch.as_tensor(np.random.RandomState(3).uniform(0.6, 1.4, (r, 2)) / 100.0).requires_grad_(True)
    f = torch.randn(n, 4, dtype=torch.float64, requires_grad=True)
    al = torch.randn(1, 2, 2, dtype=torch.float64, requires_grad=True)
    ar = torch.randn(1, 2, 2, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(
        lambda f_, al_, ar_, th_: O.regat_forward(src, dst, et, n, f_, al_, ar_, th_, 100.0, 0.2), (f, al, ar, th))
    th1 = torch.as_tensor(np.random.RandomState(4).uniform(1.2, 1.9, (r, 1)) / 100.0).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda x_, t_: O.regraphconv_forward(src, dst, et, n, x_, t_, 100.0), (f, th1))
