"""Ranked-pose output: host mirror of the writing half of ``inference.py``.

After ``sampling()`` the reference re-orders the poses by confidence and writes, per complex (inference.py:198-240):
``rank{r}.sdf`` / ``rank{r}_confidence{c:.2f}.sdf`` through RDKit (``write_mol_with_coords``,
datasets/process_mols.py:726-733) and, with flexible side chains, ``rank{r}_protein.pdb`` /
``rank{r}_confidence{c:.2f}_protein.pdb`` through Biopython (``SidechainPDBFile``, utils/visualise.py:62-132).
Once the loop itself runs at hundreds of poses per second this serial RDKit / Biopython writing dominates a
screening run, so here it is

* text-level (no RDKit / Biopython objects: the template SDF / PDB lines are kept and only the coordinate columns are
  rewritten, hydrogens dropped like ``RemoveHs`` / ``remove_hs``), and
* asynchronous (``AsyncWriter``: a small thread pool, so that complex k is written while complex k + 1 is docked).

File names, ordering and the atom <-> coordinate mapping follow the reference: ligand atom i of the graph is heavy atom i
of the SDF; a flexible residue is found by its (chain, residue number) id and its moved atoms are matched through the
sorted unique atoms of its rotatable-bond subcomponents, exactly like ``SidechainPDBFile.write``.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .inputs import RESIDUE_ATOM_ORDER, _sidechain_bonds


# ------------------------------------------------------------------------------------------- ligand
def _sdf_counts(line):
    return int(line[0:3]), int(line[3:6])


def mol_block_with_coords(template_lines, new_coords, remove_hs=True):
    """V2000 mol block of the template with the heavy atoms at ``new_coords`` ([n_heavy, 3]); hydrogens (and their bonds)
    are dropped when ``remove_hs`` (the graph was built from the H-less molecule, process_mols.py:435-454)."""
    na, nb = _sdf_counts(template_lines[3])
    atoms, bonds = template_lines[4:4 + na], template_lines[4 + na:4 + na + nb]
    is_h = [ln[31:34].strip().upper() == 'H' for ln in atoms]
    keep = [i for i in range(na) if not (remove_hs and is_h[i])]
    coords = np.asarray(new_coords, dtype=np.float64)
    heavy = [i for i in range(na) if not is_h[i]]
    if len(heavy) != len(coords):
        raise ValueError(f'template has {len(heavy)} heavy atoms, got {len(coords)} coordinates')
    xyz = {i: coords[k] for k, i in enumerate(heavy)}
    remap = {old: new + 1 for new, old in enumerate(keep)}
    out = list(template_lines[:3])
    kept_bonds = [ln for ln in bonds if int(ln[0:3]) - 1 in remap and int(ln[3:6]) - 1 in remap]
    out.append(f'{len(keep):3d}{len(kept_bonds):3d}' + template_lines[3][6:])
    for i in keep:
        ln = atoms[i]
        if i in xyz:
            x, y, z = xyz[i]
            ln = f'{x:10.4f}{y:10.4f}{z:10.4f}' + ln[30:]
        out.append(ln)
    for ln in kept_bonds:
        out.append(f'{remap[int(ln[0:3]) - 1]:3d}{remap[int(ln[3:6]) - 1]:3d}' + ln[6:])
    tail = [ln for ln in template_lines[4 + na + nb:] if not ln.startswith('M  CHG') and not ln.startswith('M  RAD')] if remove_hs and any(is_h) \
        else list(template_lines[4 + na + nb:])
    if not any(ln.startswith('M  END') for ln in tail):
        tail = ['M  END'] + tail
    out += tail[:next(i for i, ln in enumerate(tail) if ln.startswith('M  END')) + 1]
    return out


def write_mol_with_coords(template_lines, new_coords, path, remove_hs=True):
    """datasets/process_mols.py:726-733 (one-molecule SDF)."""
    with open(path, 'w') as f:
        f.write('\n'.join(mol_block_with_coords(template_lines, new_coords, remove_hs)) + '\n$$$$\n')


# ------------------------------------------------------------------------------------------- protein
def _pdb_residues(lines):
    """[(chain, resseq int, [line indices of the residue's heavy atoms in graph order])] for ATOM records."""
    res, key = [], None
    for i, ln in enumerate(lines):
        if not ln.startswith('ATOM'):
            continue
        elem = ln[76:78].strip().upper() or ln[12:16].strip()[0]
        if elem in ('H', 'D'):
            continue
        k = (ln[21], ln[22:27].strip())
        if k != key:
            key = k
            res.append([ln[21], ln[22:27].strip(), ln[17:20].strip(), []])
        res[-1][3].append(i)
    out = []
    for chain, resseq, name, idx in res:
        order = RESIDUE_ATOM_ORDER.get(name)
        if order is not None:                          # same atom order as the graph (datasets/pdbbind.py order_atoms_in_residue)
            nm = lambda i: lines[i][12:16].strip()
            idx = sorted(idx, key=lambda i: 999 if nm(i) == 'OXT' else (order.index(nm(i)) if nm(i) in order else 998))
        out.append((chain, resseq, name, idx))
    return out


def sidechain_pdb_lines(protein_lines, flex, atom_pos):
    """utils/visualise.py:62-132 for one conformation: the protein's lines with the atoms of every flexible residue moved to
    ``atom_pos`` (graph atom coordinates in the ORIGINAL frame).  ``flex``: the graph's 'flexResidues' store."""
    ids = [tuple(map(str, t)) for t in getattr(flex, 'pdbIds', [])]
    if not ids:
        return list(protein_lines)
    cum = np.concatenate([[0], np.cumsum(np.asarray(flex.residueNBondsMapping))])
    sub = np.asarray(flex.subcomponents)
    mapping = np.asarray(flex.subcomponentsMapping).reshape(-1, 2)
    lines = list(protein_lines)
    atom_pos = np.asarray(atom_pos, dtype=np.float64)
    done = set()
    for chain, resseq, name, idx in _pdb_residues(lines):
        rid = (chain, resseq)
        if rid not in ids:
            continue
        k = ids.index(rid)
        flex_atoms = np.unique(np.concatenate([sub[a:b] for a, b in mapping[cum[k]:cum[k + 1]]]))
        names = [lines[i][12:16].strip() for i in idx]
        cur_atoms = np.unique(np.concatenate([np.asarray(c) for c, _ in _sidechain_bonds(names)]))
        if len(flex_atoms) != len(cur_atoms):
            raise ValueError(f'residue {rid}: subcomponents of the PDB file do not match the flexResidues store')
        for fa, ca in zip(flex_atoms, cur_atoms):
            i = idx[int(ca)]
            x, y, z = atom_pos[int(fa)]
            lines[i] = lines[i][:30] + f'{x:8.3f}{y:8.3f}{z:8.3f}' + lines[i][54:]
        done.add(rid)
    if done != set(ids):
        raise ValueError(f'missing flexible residues in the PDB file: {set(ids) - done}')
    return lines


# ------------------------------------------------------------------------------------------- per complex
def write_ranked_poses(write_dir, result, ligand_sdf_lines, protein_pdb_lines=None, flex=None, remove_hs=True):
    """inference.py:221-240.  ``result``: what ``inference.infer_single_complex`` returns (poses already sorted by confidence).
    Returns the list of files written."""
    os.makedirs(write_dir, exist_ok=True)
    conf = result['confidence']
    files = []
    for rank, pos in enumerate(result['ligand_pos']):
        block = '\n'.join(mol_block_with_coords(ligand_sdf_lines, pos, remove_hs)) + '\n$$$$\n'
        names = ([f'rank{rank + 1}.sdf'] if rank == 0 else []) + \
            [f'rank{rank + 1}_confidence{conf[rank]:.2f}.sdf' if conf is not None else f'rank{rank + 1}_confidence.sdf']
        for nm in names:
            with open(os.path.join(write_dir, nm), 'w') as f:
                f.write(block)
            files.append(nm)
    if protein_pdb_lines is not None and flex is not None:
        for rank, apos in enumerate(result['atom_pos']):
            text = '\n'.join(sidechain_pdb_lines(protein_pdb_lines, flex, apos)) + '\n'
            names = ([f'rank{rank + 1}_protein.pdb'] if rank == 0 else []) + \
                [f'rank{rank + 1}_confidence{conf[rank]:.2f}_protein.pdb' if conf is not None else f'rank{rank + 1}_confidence_protein.pdb']
            for nm in names:
                with open(os.path.join(write_dir, nm), 'w') as f:
                    f.write(text)
                files.append(nm)
    return files


class AsyncWriter:
    """Writes finished complexes on worker threads while the next ones are docked (file formatting and I/O release the GIL
    for most of their time); ``close()`` waits for everything and re-raises the first failure."""

    def __init__(self, workers=2):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.futures = []

    def submit(self, *args, **kwargs):
        self.futures.append(self.pool.submit(write_ranked_poses, *args, **kwargs))

    def close(self):
        out = [f.result() for f in self.futures]
        self.pool.shutdown()
        return out
