"""Build the oracle models from the product models' hyper-parameters and weights (TEST INFRASTRUCTURE)."""
from functools import partial

import torch

from . import diffusion_ref as D
from .score_model_ref import TensorProductScoreModel


def oracle_model(args, state_dict, so3_score_norm, torus_score_norm, confidence_mode=False):
    """Mirror of utils/utils.py:59-108 for the oracle class; ``state_dict`` uses the reference key names."""
    has = lambda k: hasattr(args, k)
    emb = partial(D.sinusoidal_embedding, dim=args.sigma_embed_dim, scale=args.embedding_scale if has('embedding_scale') else 10000)
    m = TensorProductScoreModel(
        t_to_sigma=partial(D.t_to_sigma, args=args), timestep_emb_func=emb, so3_score_norm=so3_score_norm,
        torus_score_norm=torus_score_norm, no_torsion=args.no_torsion, num_conv_layers=args.num_conv_layers,
        lig_max_radius=args.max_radius, scale_by_sigma=args.scale_by_sigma, sh_lmax=args.sh_lmax,
        sigma_embed_dim=args.sigma_embed_dim, ns=args.ns, nv=args.nv, distance_embed_dim=args.distance_embed_dim,
        cross_distance_embed_dim=args.cross_distance_embed_dim, batch_norm=not args.no_batch_norm, dropout=args.dropout,
        use_second_order_repr=args.use_second_order_repr, cross_max_distance=args.cross_max_distance,
        dynamic_max_cross=args.dynamic_max_cross, lm_embedding_type='esm', confidence_mode=confidence_mode,
        fixed_center_conv=not args.not_fixed_center_conv if has('not_fixed_center_conv') else False,
        atom_max_neighbors=args.atom_max_neighbors, flexible_sidechains=args.flexible_sidechains,
        use_old_atom_encoder=args.use_old_atom_encoder if has('use_old_atom_encoder') else True,
        asyncronous_noise_schedule=bool(getattr(args, 'asyncronous_noise_schedule', False)))
    sd = {k: v.detach().cpu() for k, v in state_dict.items()}
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    m.eval()
    return m
