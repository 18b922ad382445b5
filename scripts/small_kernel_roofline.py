"""Algorithmic bytes / FLOPs per launch of the kernels around the fused conv, against the ncu launch list of one forward
(batch of 20 x 3dpf apo; profiles/r1_launches_forward_bf16_v23.csv: cold-cache, serialised per-launch times).
  python scripts/small_kernel_roofline.py profiles/r1_launches_forward_bf16_v23.csv > profiles/r1_small_kernels_roofline.txt
Sizes of the batch (scripts/profile_forward.py prints the edge counts): see N / E below."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = dict(lig=740, atom=21960, rec=2780)
E = dict(ll=12224, aa=175680, lr=74568, la=13493, rr=65720, ar=21960, center=740)
NS, F = 60, (90, 120, 180, 180, 180, 180)                      # f_out of the six interaction layers
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
HBM = peaks.get('hbm_gbs', 6553.0)

rows = list(csv.reader(open(sys.argv[1])))
h0 = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[h0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
launches = [(r[ik], float(r[iv]) / 1e3) for r in rows[h0 + 1:] if len(r) > iv]       # (name, us)


def pick(sub, nth):
    hits = [us for (nm, us) in launches if sub in nm]
    return hits[nth] if nth < len(hits) else None


def edge_embed_bytes(e):                                      # 2 positions + 2 indices in, hidden ns + 4 harmonics out
    return e * (24 + 8 + 4 * NS + 16)


def node_update_bytes(l):                                     # old features + one sum in, new features out, per node
    f_in = F[l - 1] if l else NS
    n = N['lig'] + N['atom'] + (N['rec'] if l < 5 else 0)
    return n * 4 * (f_in + 2 * F[l])


table = [
    ('edge_embed_fold (aa, 175680 edges)', pick('edge_embed_fold', 0), edge_embed_bytes(E['aa']), 2.0 * E['aa'] * 64 * NS),
    ('edge_embed_fold (lr, 74568 edges)', pick('edge_embed_fold', 1), edge_embed_bytes(E['lr']), 2.0 * E['lr'] * 64 * NS),
    ('edge_embed_fold (rr, 65720 edges)', pick('edge_embed_fold', 4), edge_embed_bytes(E['rr']), 2.0 * E['rr'] * 64 * NS),
    ('knn_scan_filter<9> (21960 atoms, 1098 per graph)', pick('knn_scan_filter', 0), N['atom'] * (12 + 9 * 4), None),
    ('radius_scan (lr: 740 queries x 139 residues)', pick('radius_scan', 1), N['lig'] * 139 * 12, None),
    ('degree_multi (6 edge sets)', pick('degree_multi', 0), 4 * (E['aa'] + E['lr'] * 2 + E['la'] * 2 + E['ll']), None),
] + [(f'node_update_multi (layer {l})', pick('node_update_multi', l), node_update_bytes(l), None) for l in range(6)] + [
    ('sum-arena zero fill (layer 3)', pick('FillFunctor<fl', 3), (N['lig'] + N['atom'] + N['rec']) * 4 * F[3], None),
]
print(f'kernel                                             |  time us | algorithmic MB |  GB/s | of {HBM:.0f} GB/s HBM | GFLOP/s')
for name, us, nbytes, flops in table:
    if us is None:
        continue
    gbs = nbytes / us / 1e3
    fl = f'{flops / us / 1e3:9.0f}' if flops else '        -'
    print(f'{name:50s} | {us:8.1f} | {nbytes / 1e6:14.2f} | {gbs:5.0f} | {gbs / HBM:17.3f} | {fl}')
conv = sum(us for nm, us in launches if 'tpconv_umma' in nm)
tot = sum(us for nm, us in launches)
print(f'\nsum of all {len(launches)} launches {tot:.0f} us, fused conv launches {conv:.0f} us = {conv / tot:.3f} of it')
print('Reading: every kernel here moves a few MB -- far too little to approach the HBM roof; they are bounded by launch latency,')
print('dependent-load latency (graph searches) or fp32 FMA rate (the 64 x 60 edge MLP), which is why they were fused / batched /')
print('moved to side streams instead of being tuned for bandwidth.  The HBM-heavy op of the reference, the [E, W] per-edge weight')
print('tensor (>= 80 kB per edge), does not exist here.')
