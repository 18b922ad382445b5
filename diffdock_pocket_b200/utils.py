"""Host mirror of the model factory in ``utils/utils.py`` (``get_model``, :59-113) and the argument
defaults of the published big score model / confidence model (README.md:72,88)."""
from argparse import Namespace
from functools import partial

import torch

from .all_atom_score_model import TensorProductScoreModel as AAScoreModel
from .diffusion_utils import get_timestep_embedding, t_to_sigma as t_to_sigma_compl


def get_model(args, device, t_to_sigma, no_parallel=False, confidence_mode=False):
    """Same defaulting rules as utils/utils.py:59-108 (``'x' in args`` back-compat checks included)."""
    has = lambda k: hasattr(args, k)
    if not (has('all_atoms') and args.all_atoms):
        raise NotImplementedError('only the all-atom model (--all_atoms) is on the accelerated path')
    timestep_emb_func = get_timestep_embedding(
        embedding_type=args.embedding_type if has('embedding_type') else 'sinusoidal',
        dim=args.sigma_embed_dim, scale=args.embedding_scale if has('embedding_scale') else 10000)
    model = AAScoreModel(
        t_to_sigma=t_to_sigma, device=device, no_torsion=args.no_torsion, timestep_emb_func=timestep_emb_func,
        num_conv_layers=args.num_conv_layers, lig_max_radius=args.max_radius, scale_by_sigma=args.scale_by_sigma,
        sh_lmax=args.sh_lmax, sigma_embed_dim=args.sigma_embed_dim,
        norm_by_sigma=has('norm_by_sigma') and args.norm_by_sigma, ns=args.ns, nv=args.nv,
        distance_embed_dim=args.distance_embed_dim, cross_distance_embed_dim=args.cross_distance_embed_dim,
        batch_norm=not args.no_batch_norm, dropout=args.dropout, use_second_order_repr=args.use_second_order_repr,
        cross_max_distance=args.cross_max_distance, dynamic_max_cross=args.dynamic_max_cross,
        separate_noise_schedule=args.separate_noise_schedule,
        smooth_edges=args.smooth_edges if has('smooth_edges') else False,
        odd_parity=args.odd_parity if has('odd_parity') else False, lm_embedding_type='esm',
        confidence_mode=confidence_mode,
        asyncronous_noise_schedule=args.asyncronous_noise_schedule if has('asyncronous_noise_schedule') else False,
        affinity_prediction=args.affinity_prediction if has('affinity_prediction') else False,
        parallel=args.parallel if has('parallel') else 1,
        num_confidence_outputs=len(args.rmsd_classification_cutoff) + 1
        if has('rmsd_classification_cutoff') and isinstance(args.rmsd_classification_cutoff, list) else 1,
        parallel_aggregators=args.parallel_aggregators if has('parallel_aggregators') else "",
        fixed_center_conv=not args.not_fixed_center_conv if has('not_fixed_center_conv') else False,
        no_aminoacid_identities=args.no_aminoacid_identities if has('no_aminoacid_identities') else False,
        atom_max_neighbors=args.atom_max_neighbors, flexible_sidechains=args.flexible_sidechains,
        include_miscellaneous_atoms=args.include_miscellaneous_atoms if has('include_miscellaneous_atoms') else False,
        use_old_atom_encoder=args.use_old_atom_encoder if has('use_old_atom_encoder') else True)
    model.to(device)
    return model


def score_model_args(**over):
    """model_parameters.yml equivalent of the README big score model."""
    a = dict(all_atoms=True, no_torsion=False, num_conv_layers=6, max_radius=5.0, scale_by_sigma=True, sh_lmax=1,
             sigma_embed_dim=64, ns=60, nv=10, distance_embed_dim=64, cross_distance_embed_dim=64, no_batch_norm=False,
             dropout=0.1, use_second_order_repr=False, cross_max_distance=80, dynamic_max_cross=True,
             separate_noise_schedule=False, embedding_type='sinusoidal', embedding_scale=1000, not_fixed_center_conv=False,
             atom_max_neighbors=8, flexible_sidechains=True, use_old_atom_encoder=False, c_alpha_max_neighbors=24,
             receptor_radius=15, tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.03, rot_sigma_max=1.55,
             tor_sigma_min=0.03, tor_sigma_max=3.14, sidechain_tor_sigma_min=0.03, sidechain_tor_sigma_max=3.14)
    a.update(over)
    return Namespace(**a)


def confidence_model_args(**over):
    """README.md:88 confidence model (ns=24, nv=6, 5 layers, embed dims 32, embedding_scale 10000)."""
    a = vars(score_model_args(ns=24, nv=6, num_conv_layers=5, sigma_embed_dim=32, distance_embed_dim=32,
                              cross_distance_embed_dim=32, embedding_scale=10000, dropout=0.0))
    a.update(over)
    return Namespace(**a)


def build_models(device, score_args=None, conf_args=None, seed=0, with_confidence=True, randomize_bn=True):
    """Random-init score (+ confidence) model of the published architecture: no checkpoint is available
    offline.  BatchNorm running statistics are randomised so that the folded BatchNorm is
    exercised (at init they are 0 / 1)."""
    score_args = score_args or score_model_args()
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        t2s = partial(t_to_sigma_compl, args=score_args)
        model = get_model(score_args, torch.device('cpu'), t2s, no_parallel=True)
        conf = None
        if with_confidence:
            conf_args = conf_args or confidence_model_args()
            conf = get_model(conf_args, torch.device('cpu'), partial(t_to_sigma_compl, args=conf_args), no_parallel=True,
                             confidence_mode=True)
        if randomize_bn:
            for m in [model] + ([conf] if conf is not None else []):
                for name, buf in m.named_buffers():
                    if name.endswith('running_mean'):
                        buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
                    elif name.endswith('running_var'):
                        buf.copy_(torch.rand(buf.shape, generator=g) * 0.5 + 0.75)
                for name, p in m.named_parameters():
                    if 'batch_norm' in name and name.endswith('weight'):
                        p.data.copy_(torch.rand(p.shape, generator=g) * 0.4 + 0.8)
                    elif 'batch_norm' in name and name.endswith('bias'):
                        p.data.copy_(torch.randn(p.shape, generator=g) * 0.1)
    finally:
        torch.random.set_rng_state(state)
    model.eval()
    if conf is not None:
        conf.eval()
    return model.to(device), (conf.to(device) if conf is not None else None), score_args, conf_args
