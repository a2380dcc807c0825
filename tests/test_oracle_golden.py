"""Pins the oracle (oracle/regnn_oracle.py) against the golden vectors recorded from the reference's
own layer / model code (tests/golden/make_golden.py): outputs and every gradient, float64."""
import pytest
import torch
import torch.nn.functional as F

import helpers
from oracle import regnn_oracle as O

LAYER_CASES = [c for c in helpers.golden_cases() if not c.startswith('model_')]


def _p(case, k, grad=True):
    key = 'param::' + k
    if key not in case:
        return None
    return torch.as_tensor(case[key]).requires_grad_(grad)


def oracle_layer(case, x):
    meta, kw = case['meta'], case['meta']['kw']
    src, dst = torch.as_tensor(case['src']), torch.as_tensor(case['dst'])
    et = None if meta.get('no_etype') else torch.as_tensor(case['etype'])
    n, alpha = int(case['num_nodes']), meta['alpha']
    act = helpers.ACT[kw.get('activation')]
    kind = meta['kind']
    P = {k[len('param::'):]: torch.as_tensor(v).requires_grad_(True) for k, v in case.items() if k.startswith('param::')}
    if kind == 'REGraphConv':
        out = O.regraphconv_forward(src, dst, et, n, x, P['edge_weight'], alpha, P.get('weight'), P.get('bias'), act,
                                    kw.get('norm', True), kw['in_feats'], kw['out_feats'])
    elif kind == 'RESAGEConv':
        out = O.resage_forward(src, dst, et, n, x, P['edge_weight'], alpha, P.get('weight'), P.get('bias'), act)
    elif kind == 'REGINConv':
        nrm = O.weighted_degree_norm(O.edge_relation(P['edge_weight'], alpha, et), dst, n, -1.0).unsqueeze(1)
        out = O.segment_sum(x[src] * O.edge_relation(P['edge_weight'], alpha, et), dst, n) * nrm
        out = act(out @ P['apply_func.weight'].t() + P['apply_func.bias'])
    elif kind == 'REMixHopConv':
        out = O.remixhop_forward(src, dst, et, n, x, P['edge_weight'], alpha,
                                 {j: P['weights.%d.weight' % j] for j in kw['p']}, tuple(kw['p']), act)
    elif kind == 'REGATConv':
        res_w = P.get('res_fc.weight')
        out = O.regat_forward(src, dst, et, n, x, P['attn_l'], P['attn_r'], P['edge_weight'], alpha,
                              kw.get('negative_slope', 0.2), P.get('fc.weight'), res_w,
                              residual_identity=kw.get('residual', False) and res_w is None, activation=act)
    else:
        fc_src = (P['fc_src.weight'], P.get('fc_src.bias')) if 'fc_src.weight' in P else None
        fc_dst = (P['fc_dst.weight'], P.get('fc_dst.bias')) if 'fc_dst.weight' in P else None
        if kw.get('share_weights'):
            fc_dst = fc_src
        res = (P['res_fc.weight'], P.get('res_fc.bias')) if 'res_fc.weight' in P else None
        out = O.regatv2_forward(src, dst, et, n, x, P['attn'], P['edge_weight'], alpha, kw.get('negative_slope', 0.2),
                                fc_src, fc_dst, res, residual_identity=kw.get('residual', False) and res is None,
                                activation=act, return_attention=True)
    return out, P


@pytest.mark.parametrize('name', LAYER_CASES)
def test_oracle_matches_reference_golden(name):
    case = helpers.load_case(name)
    x = torch.as_tensor(case['in::x']).requires_grad_(True)
    out, P = oracle_layer(case, x)
    outs = out if isinstance(out, tuple) else (out,)
    helpers.assert_close(outs[0].detach(), case['out0'], 1e-12, 'out')
    if 'out1' in case:
        helpers.assert_close(outs[1].detach(), case['out1'], 1e-12, 'attention')
    outs[0].backward(torch.as_tensor(case['gout']))
    helpers.assert_close(x.grad, case['gin::x'], 1e-11, 'd_x')
    shared = case['meta']['kw'].get('share_weights', False)
    for k, v in case.items():
        if not k.startswith('grad::'):
            continue
        name_ = k[len('grad::'):]
        got = P[name_].grad if P[name_].grad is not None else torch.zeros_like(P[name_])
        if shared and name_.startswith('fc_src.'):   # oracle holds the shared linear under one name only
            pass
        helpers.assert_close(got, v, 1e-11, 'd_' + name_)


def test_oracle_regcn_model_matches_reference_golden():
    for name in ('model_regcn_2layer', 'model_regcn_3layer'):
        case = helpers.load_case(name)
        meta = case['meta']
        n_layers = meta['args'][5]
        P = {k[len('param::'):]: torch.as_tensor(v) for k, v in case.items() if k.startswith('param::')}
        params = dict(fc=[(P['fc_list.%d.weight' % i], P['fc_list.%d.bias' % i]) for i in range(3)],
                      layers=[{k: P.get('layers.%d.%s' % (l, k)) for k in ('edge_weight', 'weight', 'bias')}
                              for l in range(n_layers)],
                      out=(P['out_lin.weight'], P['out_lin.bias']))
        feats = [torch.as_tensor(case['in::f%d' % i]) for i in range(3)]
        out, _ = O.regcn_model_forward(torch.as_tensor(case['src']), torch.as_tensor(case['dst']),
                                       torch.as_tensor(case['etype']), int(case['num_nodes']), feats, params,
                                       meta['args'][1], n_layers, F.elu)
        helpers.assert_close(out, case['out0'], 1e-12, name)


# ---- MAG-stack layers: fixtures from the reference's own mag/regnn_layers.py over the PyG-semantics stub ----------
@pytest.mark.parametrize('name', [c for c in helpers.mag_golden_cases() if not c.startswith('model_')])
def test_mag_oracle_matches_reference_layer_code(name):
    c = helpers.load_mag_case(name)
    m, kw = c['meta'], c['meta']['kw']
    t = lambda k: torch.as_tensor(c[k])  # noqa: E731
    p = {k[7:]: torch.as_tensor(v).clone().requires_grad_(True) for k, v in c.items() if k.startswith('param::')}
    x = t('x_src').clone().requires_grad_(True)
    n_dst = int(c['n_dst'])
    slt, res = kw.get('self_loop_type', 1), kw.get('residual', False)
    if m['kind'] == 'SaintREGCNConv':   # class lifted from the mag/regnn_saint.py script
        keep = (t('ew') != 0) if m.get('train') else None     # the mask F.dropout drew, recovered from the returned weights
        out, ew = O.saint_regcn_forward(x, t('edge_index'), t('edge_type'), p['weight'], p['bias'], p['relation_weight'],
                                        100.0, kw.get('use_softmax', False), keep, kw.get('dropout', 0.0), True)
        assert torch.allclose(ew, t('ew'), rtol=1e-11, atol=1e-14)
        common = None
    else:
        common = (x, x[:n_dst], t('edge_index'), t('edge_type'), t('target_node_type'))
    if common is None:
        pass
    elif m['kind'] == 'REGCNConv':
        out = O.mag_regcn_forward(*common, p['weight'], p['bias'], p['relation_weight'], 100.0, m['num_edge_types'], slt, res)
    elif m['kind'] == 'REGATConv':
        out = O.mag_regat_forward(*common, p['lin_src.weight'], p['att_src'], p['att_dst'], p['bias'], p['relation_weight'],
                                  100.0, m['num_edge_types'], kw['heads'], m['out_channels'],
                                  kw.get('negative_slope', 0.2), slt, res, kw.get('concat', True))
    else:
        out = O.mag_regatv2_forward(*common, p['lin_src.weight'], p['att'], p['bias'], p['relation_weight'], 100.0,
                                    m['num_edge_types'], kw['heads'], m['out_channels'], kw.get('negative_slope', 0.2),
                                    slt, res, kw.get('concat', True))
    assert torch.allclose(out, t('out'), rtol=1e-11, atol=1e-12)
    out.backward(t('gout'))
    assert torch.allclose(x.grad, t('gx_src'), rtol=1e-10, atol=1e-12)
    seen = set()
    for k, v in c.items():
        if not k.startswith('grad::'):
            continue
        key = k[6:]
        if key in ('weight_root', 'lin_dst.weight') or key not in p:    # aliases of weight / lin_src.weight
            continue
        seen.add(key)
        g = p[key].grad if p[key].grad is not None else torch.zeros_like(p[key])
        assert torch.allclose(g, torch.as_tensor(v), rtol=1e-10, atol=1e-12), key
    assert 'relation_weight' in seen and 'bias' in seen
