import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _common as T
from diffdock_pocket_b200 import inputs as inp
DEV = torch.device('cuda:0')
m, c, om, oc, sa, ca = T.models(DEV, small=True)
g1 = inp.synthetic_complex(12, n_lig=25, n_res=45, flexible_residues=3)
g2 = inp.synthetic_complex(14, n_lig=10, n_res=24, flexible_residues=0)
g3 = inp.synthetic_complex(13, n_lig=9, n_res=26, flexible_residues=1)
l1, l2, l3 = (T.randomized_list(g, 3, sa, seed=20 + i) for i, g in enumerate((g1, g2, g3)))
for name, dl in (('c1,c2,c2,c2,c3', [l1[2], l2[0], l2[1], l2[2], l3[0]]), ('c1,c3,c2', [l1[0], l3[0], l2[0]]), ('c2,c1,c3', [l2[0], l1[0], l3[0]]), ('c1,c3', [l1[0], l3[0]])):
    with torch.no_grad():
        joint = [o.cpu() for o in m(T.batch_at(dl, 0.4))]
        sep = [[o.cpu() for o in m(T.batch_at([g], 0.4))] for g in dl]
    for k, nm in enumerate(('tr', 'rot', 'tor', 'sc')):
        want = torch.cat([s[k] for s in sep])
        print(name, nm, tuple(joint[k].shape), tuple(want.shape), 'max diff %.3e' % float((joint[k] - want).abs().max()) if want.numel() else 'empty')
    with torch.no_grad():
        cj = c(T.batch_at(dl, 0.0)).cpu()
        cs = torch.cat([c(T.batch_at([g], 0.0)).cpu().reshape(-1) for g in dl])
    print(name, 'conf diff %.3e' % float((cj.reshape(-1) - cs).abs().max()))
