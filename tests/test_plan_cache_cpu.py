"""Host logic of sampling()'s resident-state cache (no GPU): the content key must be equal for deep copies of a complex,
differ as soon as ANY table a plan / pose state is built from differs, and ignore the moving coordinates."""
import copy
import os

import numpy as np
import torch

from diffdock_pocket_b200 import inputs, sampling as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _g():
    return inputs.load_graph_npz(os.path.join(GOLD, '3dpf_apo.npz'), name='3dpf_apo')


def test_key_ignores_moving_coordinates_and_storage():
    g = _g()
    k0 = S._graph_key(g, True)
    h = copy.deepcopy(g)
    h['ligand'].pos = h['ligand'].pos + 3.0                          # new start pose
    h['atom'].pos = h['atom'].pos.clone()
    h['atom'].pos[5] += 0.5                                          # a moved side-chain atom
    assert S._graph_key(h, True) == k0
    assert S._same_tables(S._graph_tables(g, True), S._graph_tables(h, True))


def test_key_sees_every_static_table():
    g = _g()
    k0 = S._graph_key(g, True)

    def changed(mutate):
        h = copy.deepcopy(g)
        mutate(h)
        same = S._same_tables(S._graph_tables(g, True), S._graph_tables(h, True))
        return S._graph_key(h, True) != k0 and not same

    def bump(store, attr, idx=0, by=1):
        def f(h):
            t = h[store] if not isinstance(store, tuple) else h[store]
            v = getattr(t, attr).clone()
            v.view(-1)[idx] += by
            setattr(t, attr, v)
        return f
    assert changed(bump('ligand', 'x'))
    assert changed(bump(('ligand', 'ligand'), 'edge_index'))
    assert changed(bump(('ligand', 'ligand'), 'edge_attr'))
    assert changed(bump('receptor', 'pos', by=0.125))
    assert changed(bump(('receptor', 'receptor'), 'edge_index'))
    assert changed(bump('atom', 'x'))
    assert changed(bump(('atom', 'receptor'), 'edge_index'))
    assert changed(bump('flexResidues', 'edge_idx'))
    assert changed(bump('flexResidues', 'subcomponents'))

    def flip_mask(h):
        m = h['ligand'].edge_mask.clone()
        m[int(torch.nonzero(~m)[0])] = True
        h['ligand'].edge_mask = m
    assert changed(flip_mask)

    def flip_rotate(h):
        mr = h['ligand'].mask_rotate
        mr = np.array(mr if isinstance(mr, np.ndarray) else mr[0]).copy()
        mr[0, 0] = not mr[0, 0]
        h['ligand'].mask_rotate = mr
    assert changed(flip_rotate)
    # without flexible side chains the flexible-residue tables are not part of the key
    assert S._graph_key(g, False) != k0
    h = copy.deepcopy(g)
    bump('flexResidues', 'edge_idx')(h)
    assert S._graph_key(h, False) == S._graph_key(g, False)


def test_receptor_language_model_features_are_sampled_next_to_exact_coordinates():
    g = _g()
    h = copy.deepcopy(g)
    x = h['receptor'].x.clone()
    x.view(-1)[::max(1, x.numel() // 251)] += 1.0                    # the strided sample of the wide ESM block
    h['receptor'].x = x
    assert S._graph_key(h, True) != S._graph_key(g, True)
