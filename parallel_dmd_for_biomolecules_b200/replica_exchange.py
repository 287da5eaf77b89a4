"""Temperature replica exchange across the replicas of one or more GPUs (new functionality -- the reference runs
its temp_0xx files sequentially by hand, qfile/script.sh:11-18).

One process per GPU.  Every exchange step all ranks all-gather (E_pot, T*) of their replicas (NCCL over NVLink when
the tensors live on the GPU, gloo in the CPU tests), then evaluate the SAME deterministic Metropolis decisions from
a shared counter RNG and apply the new temperatures to their own replicas on the device
(``dmdb_apply_temperatures``: velocities rescaled by sqrt(T_new/T_old), time constants reset, calendar rebuilt).
Temperatures are swapped, not configurations, so only 16 bytes per replica cross the wire.

Energies are in the engine's units where k_B T = setemp = 12 T* (main.F90:127; energy.f:74-78), hence
beta = 1 / (12 T*).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

#: the reference's temperature schedule (temp_018 ... temp_050)
LADDER = (0.18, 0.20, 0.22, 0.24, 0.26, 0.28, 0.30, 0.35, 0.40, 0.45, 0.50)


def _u01(seed: int, n: int) -> float:
    """counter RNG shared with the engine (splitmix64 -> 53-bit uniform); identical on every rank"""
    m = (1 << 64) - 1
    z = (seed + n * 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return (z >> 11) / 9007199254740992.0


def decide_swaps(epot: np.ndarray, tstar: np.ndarray, step: int, seed: int = 12345, ladder_size: Optional[int] = None) -> np.ndarray:
    """New T* per replica.  Replicas are grouped into ladders of ``ladder_size`` consecutive global indices
    (default: one ladder over everything); inside a ladder, temperature neighbours (k, k+1) with k of the step's
    parity attempt a swap with probability min(1, exp((beta_a - beta_b) (E_a - E_b)))."""
    epot = np.asarray(epot, dtype=np.float64)
    new_t = np.array(tstar, dtype=np.float64, copy=True)
    n = len(epot)
    L = ladder_size or n
    draw = 0
    for start in range(0, n, L):
        idx = np.arange(start, min(start + L, n))
        order = idx[np.argsort(new_t[idx], kind="stable")]  # replicas sorted by current temperature
        for k in range(step % 2, len(order) - 1, 2):
            a, b = order[k], order[k + 1]
            draw += 1
            ta, tb = new_t[a], new_t[b]
            if ta == tb:
                continue
            delta = (1.0 / (12.0 * ta) - 1.0 / (12.0 * tb)) * (epot[a] - epot[b])
            if delta >= 0 or _u01(seed + 7919 * step, start * 131 + draw) < np.exp(delta):
                new_t[a], new_t[b] = tb, ta
    return new_t


def exchange_step(dmd, step: int, seed: int = 12345, ladder_size: Optional[int] = None, group=None, device=None):
    """One exchange over all ranks of ``group`` (or a single process when torch.distributed is not initialised).
    Returns (new local temperatures, number of local replicas whose temperature changed)."""
    import torch
    import torch.distributed as dist

    epot, tstar = dmd.potential_energies()
    local = torch.from_numpy(np.stack([epot, tstar], axis=1))  # (R, 2) fp64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if device is not None:
            local = local.to(device)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local, group=group)
        allv = torch.cat(gathered, dim=0).cpu().numpy()
    else:
        rank, allv = 0, local.numpy()
    R = len(epot)
    new_all = decide_swaps(allv[:, 0], allv[:, 1], step, seed, ladder_size)
    mine = new_all[rank * R:(rank + 1) * R]
    changed = int((mine != tstar).sum())
    if changed:
        dmd.apply_temperatures(mine)
    return mine, changed
