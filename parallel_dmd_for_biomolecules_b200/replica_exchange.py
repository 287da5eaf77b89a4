"""Temperature replica exchange across the replicas of one or more GPUs (new functionality -- the reference runs
its temp_0xx files sequentially by hand, qfile/script.sh:11-18).

One process per GPU.  The product path is ``DMD.exchange`` = ``dmdb_exchange`` (include/dmdb200.h): energy
reduction, ``ncclAllGather`` of (E_pot, T*) over NVLink, the Metropolis decision (csrc/dmd_exchange.h) and the
temperature change of the local replicas all run on the device, on the library's stream; nothing but four counters
comes back to the host.  Temperatures are swapped, not configurations, so only 16 bytes per replica cross the wire.
A host with its own collective (MPI in a Fortran host; gloo in the CPU tests) gathers (E_pot, T*) itself and calls
``DMD.exchange_gathered``.

This module holds the glue (communicator set-up through ``torch.distributed``, the gloo path) and
``decide_swaps``, a pure-numpy restatement of the decision that the tests hold the device kernel against.

Energies are in the engine's units where k_B T = setemp = 12 T* (main.F90:127; energy.f:74-78), hence
beta = 1 / (12 T*).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

#: the reference's temperature schedule (temp_018 ... temp_050)
LADDER = (0.18, 0.20, 0.22, 0.24, 0.26, 0.28, 0.30, 0.35, 0.40, 0.45, 0.50)


def _u01(seed: int, n: int) -> float:
    """draw number n of the counter RNG shared with the engine (splitmix64 -> 53-bit uniform)"""
    m = (1 << 64) - 1
    z = (seed + n * 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return (z >> 11) / 9007199254740992.0


def slot_to_gathered(s: int, world: int, n_local: int) -> int:
    """ladders are cut from the slot order s = r * world + rank, so that the members of a ladder sit on different GPUs;
    the gathered arrays are rank-major (index rank * R + r)"""
    return (s % world) * n_local + s // world


def ladder_temperatures(world: int, rank: int, n_local: int, ladder=LADDER) -> np.ndarray:
    """initial T* of the local replicas: slot s holds ladder[s mod len(ladder)]; replicas beyond the last whole ladder get
    the coldest temperature (they never exchange)"""
    L, M = len(ladder), world * n_local
    t = np.empty(n_local)
    for r in range(n_local):
        s = r * world + rank
        t[r] = ladder[s % L] if s < (M // L) * L else ladder[0]
    return t


def decide_swaps(epot, tstar, step: int, seed: int = 12345, ladder_size: Optional[int] = None, world: int = 1) -> np.ndarray:
    """New T* per replica (numpy restatement of csrc/dmd_exchange.h, the checker of the device kernel).
    epot / tstar: the gathered, rank-major arrays of world * R replicas."""
    epot = np.asarray(epot, dtype=np.float64)
    new_t = np.array(tstar, dtype=np.float64, copy=True)
    M = len(epot)
    R = M // world
    L = ladder_size or min(M, 32)
    for lad in range(M // L):
        s0 = lad * L
        g = np.array([slot_to_gathered(s0 + m, world, R) for m in range(L)])
        T, E = new_t[g].copy(), epot[g]
        order = np.argsort(T, kind="stable")
        for k in range(step % 2, L - 1, 2):
            a, b = order[k], order[k + 1]
            ta, tb = T[a], T[b]
            if ta == tb:
                continue
            delta = (1.0 / (12.0 * ta) - 1.0 / (12.0 * tb)) * (E[a] - E[b])
            if delta >= 0 or _u01(seed + 7919 * step, s0 * 131 + k + 1) < np.exp(delta):
                T[a], T[b] = tb, ta
        new_t[g] = T
    return new_t


def init_communicator(dmd, group=None):
    """create the library's NCCL communicator over the ranks of torch.distributed (a no-op for one process)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        dmd.comm_init(1, 0)
        return 1, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def bcast(data):
        box = [data]
        dist.broadcast_object_list(box, src=0, group=group)
        return box[0]

    dmd.comm_init(world, rank, bcast)
    return world, rank


def exchange_step(dmd, step: int, seed: int = 12345, ladder_size: Optional[int] = None, group=None):
    """One exchange through a HOST-side gather (gloo / any torch.distributed backend): all ranks all-gather
    (E_pot, T*), the decision and the temperature change are made by ``dmdb_exchange_gathered``.  The GPU path
    with the collective on the device is ``dmd.exchange`` (see bench.py).  Returns the dmdb_exchange_stats."""
    import torch
    import torch.distributed as dist

    epot, tstar = dmd.potential_energies()
    local = torch.from_numpy(np.stack([epot, tstar], axis=1))  # (R, 2) fp64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local, group=group)
        allv = torch.cat(gathered, dim=0).numpy()
    else:
        world, rank, allv = 1, 0, local.numpy()
    return dmd.exchange_gathered(allv, world, rank, step, seed, ladder_size or 0)
