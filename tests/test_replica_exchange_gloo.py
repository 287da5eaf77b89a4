"""Multi-process path of the replica exchange with world_size 2 over gloo on the CPU.  Each rank holds a real engine
handle -- the 1-lane host-trace build of the engine source, which runs the SAME decision code (csrc/dmd_exchange.h) as
the device kernel -- gathers (E_pot, T*) over gloo and calls dmdb_exchange_gathered.  Both ranks must take the
decisions of the numpy restatement, the multiset of temperatures of every ladder must be conserved (ladders are cut
across the ranks), and detailed balance must hold for the acceptance rule."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import HOSTTRACE
from parallel_dmd_for_biomolecules_b200 import genconfig, tables
from parallel_dmd_for_biomolecules_b200 import replica_exchange as rx
from parallel_dmd_for_biomolecules_b200.dmd import DMD

R_LOCAL, STEPS = 11, 4


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = tables.load_default_tables()
    topo, sv = genconfig.generate_box(["AAAA"], [3], 40.0, 0.3, tab, seed=2)
    p = tables.make_params(boxl=40.0, tstar=0.18, canon=True, n_replicas=R_LOCAL, seed=100 + 1000 * rank)
    d = DMD(p, topo, tab, lib_path=HOSTTRACE)
    d.set_state(sv)
    d.apply_temperatures(rx.ladder_temperatures(world, rank, R_LOCAL))
    hist = []
    for step in range(STEPS):
        d.run(4000)
        epot, before = d.potential_energies()
        st = rx.exchange_step(d, step, seed=42, ladder_size=11)
        after = d.potential_energies()[1]
        assert st.changed_local == int((after != before).sum())
        hist.append((epot, before, after, (st.ladders, st.attempted, st.accepted)))
    out[rank] = hist
    d.close()
    dist.destroy_process_group()


def test_two_rank_exchange_is_consistent(hosttrace_lib):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    swapped = 0
    for step in range(STEPS):
        (e0, b0, a0, c0), (e1, b1, a1, c1) = out[0][step], out[1][step]
        assert c0 == c1 and c0[0] == 2  # both ranks counted the same attempts / acceptances over two ladders
        t = rx.decide_swaps(np.concatenate([e0, e1]), np.concatenate([b0, b1]), step, seed=42, ladder_size=11, world=2)
        assert np.array_equal(t[:R_LOCAL], a0) and np.array_equal(t[R_LOCAL:], a1)  # the numpy restatement agrees
        for lad in range(2):  # a ladder's members alternate between the ranks; its temperatures are conserved
            g = [rx.slot_to_gathered(lad * 11 + m, 2, R_LOCAL) for m in range(11)]
            assert sorted(t[g]) == sorted(rx.LADDER)
            assert {x // R_LOCAL for x in g} == {0, 1}
        swapped += c0[2]
    assert swapped > 0  # something was exchanged


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


def test_engine_decision_equals_numpy_restatement(hosttrace_lib, tab):
    """dmdb_exchange_gathered on arbitrary gathered arrays (several ranks' worth, ragged tail outside the last ladder)"""
    topo, sv = genconfig.generate_box(["AAAA"], [2], 40.0, 0.3, tab, seed=2)
    R, world = 7, 5  # 35 replicas = 3 ladders of 11 + 2 left over
    rng = np.random.default_rng(3)
    for rank in (0, 3):
        d = DMD(tables.make_params(boxl=40.0, tstar=0.3, canon=True, n_replicas=R, seed=5), topo, tab, lib_path=HOSTTRACE)
        d.set_state(sv)
        tl = rx.ladder_temperatures(world, rank, R)
        d.apply_temperatures(tl)
        for step in range(3):
            e = rng.normal(-40.0, 25.0, world * R)
            t = np.concatenate([rx.ladder_temperatures(world, q, R) for q in range(world)])
            rng.shuffle(t.reshape(-1)[: (world * R // 11) * 11 // 1])  # any assignment of temperatures to replicas
            t[rank * R:(rank + 1) * R] = d.potential_energies()[1]      # ... consistent with the local handle
            st = d.exchange_gathered(np.stack([e, t], axis=1), world, rank, step, seed=9, ladder_size=11)
            want = rx.decide_swaps(e, t, step, seed=9, ladder_size=11, world=world)
            assert np.array_equal(d.potential_energies()[1], want[rank * R:(rank + 1) * R])
            assert st.ladders == 3 and st.attempted == 3 * (5 if step % 2 == 0 else 5)
        d.close()
