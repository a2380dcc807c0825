"""Shared test helpers: golden-fixture loading and module construction from fixture metadata."""
import glob
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ACT = {'elu': F.elu, 'relu': F.relu, None: None}


def golden_cases(prefix=''):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    case = {k: z[k] for k in z.files}
    case['meta'] = json.loads(str(case['meta']))
    return case


def build_layer(pkg, meta):
    """Instantiates ``pkg.<kind>`` (our package or the reference's ``layer``) from fixture metadata."""
    kw = dict(meta['kw'])
    if 'activation' in kw:
        kw['activation'] = ACT[kw['activation']]
    if 'apply_linear' in meta:
        kw['apply_func'] = torch.nn.Linear(*meta['apply_linear'])
    return getattr(pkg, meta['kind'])(meta['R'], meta['alpha'], **kw)


def load_params(module, case, dtype=torch.float64, device='cpu'):
    module = module.to(dtype=dtype, device=device)
    sd = {k[len('param::'):]: torch.as_tensor(v) for k, v in case.items() if k.startswith('param::')}
    missing, unexpected = module.load_state_dict({k: v.to(dtype=dtype, device=device) for k, v in sd.items()},
                                                 strict=True)
    assert not missing and not unexpected
    return module


def f32_exact(a):
    """float64 tensor whose values are exactly representable in float32, so the float64 oracle and
    the fp32 kernels start from identical inputs (and LeakyReLU kinks are hit on the same side)."""
    return torch.as_tensor(np.asarray(a, dtype=np.float32).astype(np.float64))


def assert_close(actual, expected, rtol, name='', max_outlier_frac=0.0, atol=0.0):
    """Norm-wise relative check used for fp32 results: |a-e| <= rtol * (|e| + max|e|) elementwise.
    (Sums with cancellation make a purely elementwise relative bound meaningless near zero.)
    ``max_outlier_frac``: fraction of elements allowed outside the bound -- only used downstream of a
    LeakyReLU kink fed by an fp32 GEMM, where a pre-activation within 1 ulp of zero can land on the
    other side of the kink than in float64 and flips one gradient term (a property of fp32, also of
    the reference's own fp32 run)."""
    a = torch.as_tensor(np.asarray(actual), dtype=torch.float64)
    e = torch.as_tensor(np.asarray(expected), dtype=torch.float64)
    assert a.shape == e.shape, '%s: shape %s vs %s' % (name, tuple(a.shape), tuple(e.shape))
    assert torch.isfinite(a).all(), '%s: non-finite values' % name
    scale = e.abs().max().item() if e.numel() else 0.0
    err = (a - e).abs()
    bound = rtol * (e.abs() + scale) + 1e-30 + atol   # atol: for quantities whose exact value is 0 (pure rounding noise)
    ratio = err / bound
    if max_outlier_frac > 0 and e.numel():
        k = int(max_outlier_frac * e.numel())
        if k > 0:
            ratio = torch.topk(ratio.reshape(-1), k + 1, largest=True).values[-1:]
    worst = ratio.max().item() if e.numel() else 0.0
    # measured errors, printed for passes too (``pytest -rP`` shows them): the norm-wise relative error
    # max|a-e| / max|e| and the largest purely elementwise relative error over elements with |e| >= 1e-3 * max|e|
    big = e.abs() >= 1e-3 * scale if scale > 0 else torch.zeros_like(e, dtype=torch.bool)
    elem = (err[big] / e.abs()[big]).max().item() if bool(big.any()) else 0.0
    print('[parity] %-28s n=%-9d max|err|/max|ref| = %.2e   max elementwise rel (|ref| >= 1e-3 max) = %.2e   '
          'bound use %.2f of rtol=%g' % (name, e.numel(), (err.max().item() / scale) if scale > 0 else 0.0, elem, worst, rtol))
    assert worst <= 1.0, '%s: max err %.3e (scale %.3e), %.2fx over rtol=%g' % (
        name, err.max().item(), scale, worst, rtol)


def build_model(model_pkg, meta, g):
    """Instantiates REGCN / REGAT / REMixHop (ours or the reference's) from fixture metadata."""
    args = [ACT.get(a, a) if isinstance(a, str) else a for a in meta['args']]
    if meta['kind'] == 'REGCN':
        return model_pkg.REGCN(g, *args)
    if meta['kind'] == 'REGAT':
        return model_pkg.REGAT(g, *args, use_gatv2=meta.get('use_gatv2', False))
    if meta['kind'] == 'REGIN':
        return model_pkg.REGIN(g, *args)
    return model_pkg.REMixHop(g, *args, activation=ACT[meta.get('activation')])


def run_model_case(model_pkg, graph_cls, case, device='cpu', dtype=torch.float64):
    g = graph_cls(case['src'], case['dst'], int(case['num_nodes'])).to(device)
    net = load_params(build_model(model_pkg, case['meta'], g), case, dtype, device)
    feats = [torch.as_tensor(case['in::f%d' % i]).to(device=device, dtype=dtype).requires_grad_(True) for i in range(3)]
    out, _ = net(feats, torch.as_tensor(case['etype']).to(device))
    out.backward(torch.as_tensor(case['gout']).to(device=device, dtype=dtype))
    return net, feats, out


def check_model_case(case, net, feats, out, rtol, grad_rtol=None):
    grad_rtol = grad_rtol or rtol
    assert_close(out.detach().cpu(), case['out0'], rtol, 'logits')
    for i, f in enumerate(feats):
        assert_close(f.grad.cpu(), case['gin::f%d' % i], grad_rtol, 'd_f%d' % i)
    for k, p in net.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert_close(got.cpu(), case['grad::' + k], grad_rtol, 'd_' + k)


MAG_GOLDEN = os.path.join(GOLDEN, 'mag')


def mag_golden_cases(prefix=''):
    """Fixtures recorded from the reference's own mag/regnn_layers.py over oracle/pyg_stub (make_golden_mag.py)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(MAG_GOLDEN, prefix + '*.npz')))


def load_mag_case(name):
    z = np.load(os.path.join(MAG_GOLDEN, name + '.npz'), allow_pickle=False)
    case = {k: z[k] for k in z.files}
    case['meta'] = json.loads(str(case['meta']))
    return case
