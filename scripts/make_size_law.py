"""Heavy-atom counts of the PDBBind test ligands (BASELINE.json configs[3] "testset_csv size distribution"), parsed from the
reference's ``data/test_ligands_smiles.txt`` with a small SMILES tokenizer (RDKit is not installed) and stored as
``tests/golden/pdbbind_test_ligand_sizes.txt`` so that the synthetic set travels to machines without the reference tree.
Run HERE (needs /root/reference)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/data/test_ligands_smiles.txt'
ATOM = re.compile(r'\[([^\]]+)\]|Cl|Br|[BCNOPSFI]|[bcnops]')


def heavy_atoms(smiles):
    n = 0
    for m in ATOM.finditer(smiles):
        if m.group(1) is not None:
            sym = re.match(r'\d*([A-Z][a-z]?|[a-z]{1,2})', m.group(1)).group(1)
            if sym == 'H':
                continue
        n += 1
    return n


sizes = [heavy_atoms(l.strip()) for l in open(SRC) if l.strip()]
out = os.path.join(ROOT, 'tests', 'golden', 'pdbbind_test_ligand_sizes.txt')
open(out, 'w').write('\n'.join(map(str, sizes)) + '\n')
s = sorted(sizes)
q = lambda p: s[int(p * (len(s) - 1))]
print(len(sizes), 'ligands: min', s[0], 'p10', q(.1), 'p25', q(.25), 'median', q(.5), 'mean', round(sum(s) / len(s), 1), 'p75', q(.75), 'p90', q(.9), 'max', s[-1])
