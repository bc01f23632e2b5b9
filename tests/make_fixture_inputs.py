"""Seeded initial fields shared by oracle/make_reference_fixtures.py and the live reference-source test."""
from helpers import O


def initial_fields(p, seed=41, amp=0.3):
    u, v, w = O.synthetic_global(p.nx, p.ny, p.Nz, nproc=1, seed=seed, amp=amp, L_x=p.L_x, L_y=p.L_y, L_z=p.L_z)
    return tuple(O.scatter_slab(f, p) for f in (u, v, w))
