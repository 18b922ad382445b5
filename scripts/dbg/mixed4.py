import copy, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _common as T
from diffdock_pocket_b200 import inputs as inp
DEV = torch.device('cuda:0')
m, c, om, oc, sa, ca = T.models(DEV, small=True)
print('conv_mode', m.conv_mode)
g1 = inp.synthetic_complex(12, n_lig=25, n_res=45, flexible_residues=3)
g3 = inp.synthetic_complex(13, n_lig=9, n_res=26, flexible_residues=1)
l1 = T.randomized_list(g1, 3, sa, seed=20)
l3 = T.randomized_list(g3, 3, sa, seed=22)
dl = [l1[0], l3[0]]
with torch.no_grad():
    want = [[o.clone() for o in om(T.oracle_batch_at([g], 0.4))] for g in dl]
    for rep in range(2):
        joint = [o.cpu().clone() for o in m(T.batch_at(dl, 0.4))]
        sep = [[o.cpu().clone() for o in m(T.batch_at([g], 0.4))] for g in dl]
        bt = T.batch_at(dl, 0.4)
        pl = m.make_plan(copy.deepcopy(bt))
        jp = [o.cpu().clone() for o in m.run_plan(pl, bt.complex_t)]
        for k, nm in enumerate(('tr', 'rot', 'tor', 'sc')):
            w = torch.cat([x[k] for x in want]); s = torch.cat([x[k] for x in sep])
            print(rep, nm, 'joint-forward vs oracle %.2e | sep-forward vs oracle %.2e | joint-run_plan vs oracle %.2e' % (
                float((joint[k] - w).abs().max()), float((s - w).abs().max()), float((jp[k] - w).abs().max())))
