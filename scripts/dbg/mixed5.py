import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _common as T
from diffdock_pocket_b200 import inputs as inp
DEV = torch.device('cuda:0')
m, c, om, oc, sa, ca = T.models(DEV, small=True)
g1 = inp.synthetic_complex(12, n_lig=25, n_res=45, flexible_residues=3)
g3 = inp.synthetic_complex(13, n_lig=9, n_res=26, flexible_residues=1)
l1 = T.randomized_list(g1, 3, sa, seed=20)
l3 = T.randomized_list(g3, 3, sa, seed=22)
dl = [l1[0], l3[0]]
with torch.no_grad():
    want = om(T.oracle_batch_at(dl, 0.4))
    dbg = om._debug
    bt = T.batch_at(dl, 0.4)
    pl = m.make_plan(copy.deepcopy(bt))
    got = [o.cpu().clone() for o in m.run_plan(pl, bt.complex_t, return_layers=True)]
for nm in ('ll', 'aa', 'lr', 'la'):
    a, b = pl.es[nm].edge_index().cpu(), dbg[nm].long()
    same = a.shape == b.shape and bool(torch.equal(a, b))
    print(nm, tuple(a.shape), tuple(b.shape), 'equal' if same else 'DIFFERENT')
    if not same and a.shape == b.shape:
        bad = (a != b).any(0).nonzero().flatten()
        print('  first mismatches at', bad[:10].tolist(), a[:, bad[:5]].tolist(), b[:, bad[:5]].tolist())
    elif not same:
        sa_ = set(map(tuple, a.T.tolist())); sb_ = set(map(tuple, b.T.tolist()))
        print('  only in product', sorted(sa_ - sb_)[:10], ' only in oracle', sorted(sb_ - sa_)[:10])
for l, ((gl, ga, gr), (wl, wa, wr)) in enumerate(zip(pl.last_layers, dbg['layers'])):
    print('layer', l, 'lig %.2e atom %.2e rec %.2e' % (T.rel_err(gl, wl), T.rel_err(ga[:, :wa.shape[1]], wa), T.rel_err(gr[:, :wr.shape[1]], wr)))
for k, nm in enumerate(('tr', 'rot', 'tor', 'sc')):
    print(nm, '%.2e' % float((got[k] - want[k]).abs().max()))
print('sizes NL', pl.NL, 'NA', pl.NA, 'NR', pl.NR, 'lig_ptr', pl.lig_ptr_h.tolist(), 'atom_ptr', pl.atom_ptr_h.tolist(), 'rec_ptr', pl.rec_ptr_h.tolist())
