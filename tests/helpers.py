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


def assert_close(actual, expected, rtol, name=''):
    """Norm-wise relative check used for fp32 results: |a-e| <= rtol * (|e| + max|e|) elementwise.
    (Sums with cancellation make a purely elementwise relative bound meaningless near zero.)"""
    a = torch.as_tensor(np.asarray(actual), dtype=torch.float64)
    e = torch.as_tensor(np.asarray(expected), dtype=torch.float64)
    assert a.shape == e.shape, '%s: shape %s vs %s' % (name, tuple(a.shape), tuple(e.shape))
    assert torch.isfinite(a).all(), '%s: non-finite values' % name
    scale = e.abs().max().item() if e.numel() else 0.0
    err = (a - e).abs()
    bound = rtol * (e.abs() + scale) + 1e-30
    worst = (err / bound).max().item() if e.numel() else 0.0
    assert worst <= 1.0, '%s: max err %.3e (scale %.3e), %.2fx over rtol=%g' % (
        name, err.max().item(), scale, worst, rtol)
