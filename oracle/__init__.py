"""CPU oracle for the DiffDock-Pocket reverse-diffusion score-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``diffdock_pocket_b200``) imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` do, and there only as the checker / the CPU arm.

It is a plain PyTorch-FP32 (CPU) + C (gcc, ``cluster_c.c``) restatement of the
reference algorithm; every function cites the reference ``file:line`` it
follows (paths relative to ``/root/reference``).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and cannot be imported in this container (e3nn,
torch_geometric, torch_cluster, torch_scatter, rdkit ... are not installed,
SURVEY.md F4).  The arithmetic that lives in those un-vendored dependencies
(e3nn 0.5.1, pytorch-cluster 1.6.1, pytorch-scatter 2.1.0, pyg 2.4.0) is
restated from their published algorithms and pinned only by
algebraic self-checks (tests/test_oracle_*.py): SO(3) equivariance, CG
identities, ``FasterTensorProduct`` == generic Clebsch-Gordan FCTP at lmax=1,
brute-force neighbour search, scipy Rotation.
"""
