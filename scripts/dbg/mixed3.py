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
a = T.randomized_list(g1, 2, sa, seed=21)
b = T.randomized_list(g3, 2, sa, seed=22)
def fwd(dl, t=0.4):
    with torch.no_grad():
        bt = T.batch_at(dl, t)
        pl = m.make_plan(copy.deepcopy(bt))
        out = [o.cpu().clone() for o in m.run_plan(pl, bt.complex_t, return_layers=True)]
        lay = [[x.cpu().clone() for x in L] for L in pl.last_layers]
        ne = {k: int(pl.es[k].n_dev.item()) for k in ('ll', 'lr', 'la', 'aa')}
    return out, lay, ne, pl
ref_a, lay_a, ne_a, _ = fwd([a[0]])
ref_b, lay_b, ne_b, _ = fwd([b[0]])
with torch.no_grad():
    wa = om(T.oracle_batch_at([a[0]], 0.4)); wb = om(T.oracle_batch_at([b[0]], 0.4))
print('alone vs oracle: a tr %.2e, b tr %.2e' % (float((ref_a[0] - wa[0]).abs().max()), float((ref_b[0] - wb[0]).abs().max())))
for name, dl, ia, ib in (('a,a2', [a[0], a[1]], 0, None), ('a,b', [a[0], b[0]], 0, 1), ('b,a', [b[0], a[0]], 1, 0), ('b,b2', [b[0], b[1]], None, 0)):
    out, lay, ne, pl = fwd(dl)
    msg = [name, str(ne)]
    nl = [g['ligand'].pos.shape[0] for g in dl]; na = [g['atom'].pos.shape[0] for g in dl]; nr = [g['receptor'].pos.shape[0] for g in dl]
    for idx, ref, rl, tag in ((ia, ref_a, lay_a, 'a'), (ib, ref_b, lay_b, 'b')):
        if idx is None:
            continue
        msg.append('%s: tr diff %.2e rot %.2e' % (tag, float((out[0][idx] - ref[0][0]).abs().max()), float((out[1][idx] - ref[1][0]).abs().max())))
        lo_l, lo_a, lo_r = sum(nl[:idx]), sum(na[:idx]), sum(nr[:idx])
        for l in range(len(lay)):
            dl_ = float((lay[l][0][lo_l:lo_l + nl[idx]] - rl[l][0]).abs().max())
            da_ = float((lay[l][1][lo_a:lo_a + na[idx]] - rl[l][1]).abs().max())
            dr_ = float((lay[l][2][lo_r:lo_r + nr[idx]] - rl[l][2]).abs().max())
            msg.append('L%d lig %.1e atom %.1e rec %.1e' % (l, dl_, da_, dr_))
    print(' | '.join(msg))
print('alone edges a', ne_a, 'b', ne_b)
