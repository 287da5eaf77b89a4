"""Multi-process path of the replica exchange with world_size 2 over gloo on the CPU: both ranks must take the
same decisions, the union of temperatures must be conserved, and detailed balance must hold for the
acceptance rule."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parallel_dmd_for_biomolecules_b200 import replica_exchange as rx


class FakeDMD:
    """stands in for the device handle: the exchange only needs potential_energies / apply_temperatures"""

    def __init__(self, epot, tstar):
        self.epot, self.tstar, self.applied = np.array(epot, float), np.array(tstar, float), None

    def potential_energies(self):
        return self.epot, self.tstar

    def apply_temperatures(self, t):
        self.applied = np.array(t)
        self.tstar = np.array(t)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    R = 11
    d = FakeDMD(rng.normal(-50, 30, R), rx.LADDER)
    hist = []
    for step in range(6):
        new_t, changed = rx.exchange_step(d, step, seed=42, ladder_size=11)
        hist.append(new_t.copy())
    out[rank] = (np.array(hist), d.epot)
    dist.destroy_process_group()


def test_two_rank_exchange_is_consistent():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    h0, e0 = out[0]
    h1, e1 = out[1]
    # replay the decisions in one process: identical
    epot = np.concatenate([e0, e1])
    t = np.concatenate([np.array(rx.LADDER), np.array(rx.LADDER)])
    for step in range(6):
        t = rx.decide_swaps(epot, t, step, seed=42, ladder_size=11)
        assert np.array_equal(t[:11], h0[step]) and np.array_equal(t[11:], h1[step])
        assert sorted(t[:11]) == sorted(rx.LADDER) and sorted(t[11:]) == sorted(rx.LADDER)
    assert not np.array_equal(h0[-1], np.array(rx.LADDER))  # something was exchanged


def test_acceptance_rule():
    # the colder replica has the higher energy: always accepted
    t = rx.decide_swaps([10.0, -10.0], [0.18, 0.20], step=0)
    assert list(t) == [0.20, 0.18]
    # hot replica much higher in energy: essentially never accepted
    acc = sum(rx.decide_swaps([-500.0, 500.0], [0.18, 0.50], step=0, seed=s)[0] != 0.18 for s in range(200))
    assert acc == 0
    # empirical acceptance matches exp(delta)
    ea, eb, ta, tb = -3.0, 2.0, 0.18, 0.20
    delta = (1 / (12 * ta) - 1 / (12 * tb)) * (ea - eb)
    acc = np.mean([rx.decide_swaps([ea, eb], [ta, tb], step=0, seed=s)[0] != ta for s in range(4000)])
    assert abs(acc - np.exp(delta)) < 0.03
    # odd steps pair (1,2): with two replicas nothing happens
    assert list(rx.decide_swaps([10.0, -10.0], [0.18, 0.20], step=1)) == [0.18, 0.20]
