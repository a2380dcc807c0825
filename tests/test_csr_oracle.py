"""The integer oracle's own invariants (it is the bit-exact yardstick for regnn_csr_build)."""
import numpy as np

from oracle import csr_oracle
from re_gnn_b200 import synth


def test_csr_oracle_invariants():
    d = synth.random_multigraph(200, 3000, 6, seed=3)
    c = csr_oracle.csr_build(d['src'], d['dst'], 200)
    e = d['src'].size
    assert c['indptr'][0] == 0 and c['indptr'][-1] == e and np.all(np.diff(c['indptr']) >= 0)
    assert np.array_equal(np.sort(c['eid']), np.arange(e))
    assert np.array_equal(d['dst'][c['eid']], c['row']) and np.array_equal(d['src'][c['eid']], c['indices'])
    for v in (0, 1, 7, 199):   # slots of a row are in increasing edge id (stable)
        s = c['eid'][c['indptr'][v]:c['indptr'][v + 1]]
        assert np.all(np.diff(s) > 0)
    # transposed view: entry j is edge (row_t -> indices_t[j]) stored at CSR slot slot_t[j]
    row_t = np.repeat(np.arange(200), np.diff(c['indptr_t']))
    assert np.array_equal(c['indices'][c['slot_t']], row_t)
    assert np.array_equal(c['row'][c['slot_t']], c['indices_t'])
    a, b = csr_oracle.etype_permute(d['etype'], c['eid'], c['slot_t'])
    assert a.dtype == np.uint8 and np.array_equal(a.astype(np.int64) + 1, d['etype'][c['eid']])
    assert np.array_equal(b, a[c['slot_t']])


def test_empty_graph():
    c = csr_oracle.csr_build(np.zeros(0, np.int64), np.zeros(0, np.int64), 4)
    assert c['indptr'].tolist() == [0] * 5 and c['indices'].size == 0


def test_named_shapes_have_the_published_sizes():
    for name, n, e, r in [('dblp', 26128, 265694, 10), ('acm', 10942, 558814, 12), ('imdb', 21420, 108062, 10)]:
        d = synth.hetero_graph(name)
        assert d['num_nodes'] == n and d['num_relations'] == r
        assert abs(d['src'].size - e) <= 16          # a handful of P-P self pairs are dropped
        assert d['etype'].min() == 1 and d['etype'].max() == r
        # one self loop per node, appended last, typed num_etype + 1 + ntype
        assert np.array_equal(d['src'][-n:], np.arange(n)) and np.array_equal(d['dst'][-n:], np.arange(n))
        assert np.array_equal(d['etype'][-n:], d['num_etype'] + 1 + d['ntype'])
