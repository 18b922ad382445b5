import copy, os, sys
from functools import partial
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _common as T
from diffdock_pocket_b200 import diffusion_utils as du, inputs as inp, sampling as ps
DEV = torch.device('cuda:0')
m, c, om, oc, sa, ca = T.models(DEV, small=True)
for variant in ('three', 'four'):
    graphs = [inp.synthetic_complex(11, n_lig=12, n_res=30, flexible_residues=2), inp.synthetic_complex(12, n_lig=25, n_res=45, flexible_residues=3)]
    if variant == 'four':
        graphs.append(inp.synthetic_complex(14, n_lig=10, n_res=24, flexible_residues=0))
    graphs.append(inp.synthetic_complex(13, n_lig=9, n_res=26, flexible_residues=1))
    lists = [T.randomized_list(g, 3, sa, seed=20 + i) for i, g in enumerate(graphs)]
    steps = 5
    sch = du.get_t_schedule('expbeta', steps)
    kw = dict(temp_sampling=[0.9766, 6.0774, 6.7616, 1.4488], temp_psi=[1.5103, 0.8141, 0.7662, 1.3396], temp_sigma_data=0.48884,
              no_random=True, confidence_model=c, filtering_model_args=ca)
    for conc in (True, False):
        run = lambda dl, bs: ps.sampling(copy.deepcopy(dl), m, steps, sch, sch, sch, sch, DEV, partial(du.t_to_sigma, args=sa), sa, batch_size=bs,
                                         concurrent_batches=conc, **kw)
        joint, conf_j = run([g for dl in lists for g in dl], 5)
        k = 0
        for ci, dl in enumerate(lists):
            sep, conf_s = run(dl, 3)
            d = max(float((a['ligand'].pos - b['ligand'].pos).abs().max()) for a, b in zip(joint[k:k + 3], sep))
            da = max(float((a['atom'].pos - b['atom'].pos).abs().max()) for a, b in zip(joint[k:k + 3], sep))
            print(variant, 'concurrent' if conc else 'single', 'complex', ci, 'lig diff %.4f atom diff %.4f conf diff %.2e' % (d, da, float((conf_j[k:k+3].cpu() - conf_s.cpu()).abs().max())))
            k += 3
