"""β-sheet / fibril observables (observables.py = results/r/fibril_list_assign.f definitions) on constructed states."""
import numpy as np

from parallel_dmd_for_biomolecules_b200 import genconfig, observables, tables


def test_sheets_from_constructed_hydrogen_bonds(tab):
    topo, sv = genconfig.system_b(tab, 0.18, seed=1, n_chains=6)
    N, L, nbd = topo.n_beads, 7, 28
    bp = np.zeros(N, dtype=np.int32)

    def bond(chain_a, res_n, chain_b, res_c):  # N of residue res_n of a  <->  C of residue res_c of b (1-based)
        n = chain_a * nbd + L + (res_n - 1)
        c = chain_b * nbd + 2 * L + (res_c - 1)
        bp[n], bp[c] = c + 1, n + 1

    # chains 0-1 share 4 H-bonds (>= 7//2+1 = 4): a sheet; chains 1-2 share 4: same sheet; chains 3-4 only 3: none
    for r in (2, 3, 4, 5):
        bond(0, r, 1, r)  # N of chain 0 with C of chain 1
        bond(1, r, 2, r)  # N of chain 1 with C of chain 2
    for r in (2, 3, 4):
        bond(3, r, 4, r)
    res = observables.sheets_and_fibrils(topo, tab, sv[:, :3], bp, 158.54)
    assert res["hb_contact"][0, 1] == 4 and res["hb_contact"][1, 2] == 4 and res["hb_contact"][3, 4] == 3
    assert res["sheets"] == [[0, 1, 2]] and res["peptides_in_sheets"] == 3 and res["fibrils"] == []
    # terminal beads do not count (fibril_list_assign.f:51): N of residue 1
    bp2 = np.zeros(N, dtype=np.int32)
    n, c = 0 * nbd + L, 1 * nbd + 2 * L + 2
    bp2[n], bp2[c] = c + 1, n + 1
    assert observables.contacts(topo, tab, sv[:, :3], bp2, 158.54)[0].sum() == 0


def test_hydrophobic_contacts_of_a_dilute_box(tab):
    topo, sv = genconfig.system_b(tab, 0.18, seed=1)
    hb, hp = observables.contacts(topo, tab, sv[:, :3], np.zeros(topo.n_beads, dtype=np.int32), 158.54)
    # chains are placed >= 5 A apart (gen_config_random-SQZ.f90:665-694), wells reach 5.5-6.9 A: a handful of contacts
    assert hb.sum() == 0 and np.array_equal(hp, hp.T) and hp.sum() < 40 and np.all(np.diag(hp) == 0)
    assert observables.sheets_and_fibrils(topo, tab, sv[:, :3], np.zeros(topo.n_beads, dtype=np.int32), 158.54)["sheets"] == []


def check_device_sheet_observables(tab, lib_path, n_rep=3):
    """dmdb_sheet_observables (the engine's own reduction over resident state) against the numpy restatement, on states
    with constructed hydrogen-bond patterns: two sheets of different size, a pair one bond short, terminal-bead bonds
    that must not count, an intra-chain bond"""
    from parallel_dmd_for_biomolecules_b200.dmd import DMD
    topo, sv = genconfig.system_b(tab, 0.18, seed=1, n_chains=10)
    N, L, nbd = topo.n_beads, 7, 28
    rng = np.random.default_rng(4)
    d = DMD(tables.make_params(boxl=158.54, tstar=0.18, canon=True, n_replicas=n_rep), topo, tab, lib_path=lib_path)
    want = []
    for rep in range(n_rep):
        bp = np.zeros(N, dtype=np.int32)

        def bond(chain_a, res_n, chain_b, res_c):
            n = chain_a * nbd + L + (res_n - 1)
            c = chain_b * nbd + 2 * L + (res_c - 1)
            if bp[n] == 0 and bp[c] == 0:
                bp[n], bp[c] = c + 1, n + 1

        order = rng.permutation(10)
        for a, b in zip(order[:3], order[1:4]):      # a sheet of four peptides (three partner pairs)
            for r in (2, 3, 4, 5):
                bond(a, r, b, r)
        for r in (2, 3, 4, 5, 6):                    # a sheet of two
            bond(order[5], r, order[6], r)
        for r in (2, 3, 4):                          # one bond short of a sheet partner
            bond(order[7], r, order[8], r)
        bond(order[9], 1, order[8], 7)               # first N with last C: bonds of end beads do not count
        bond(order[4], 6, order[4], 2)               # intra-chain
        d.set_state(sv, bp, replica=rep)
        res = observables.sheets_and_fibrils(topo, tab, sv[:, :3], bp, 158.54)
        inter = sum(1 for k in range(N) if bp[k] > k + 1 and (bp[k] - 1) // nbd != k // nbd)
        intra = sum(1 for k in range(N) if bp[k] > k + 1 and (bp[k] - 1) // nbd == k // nbd)
        dimers = int(((res["hb_contact"] >= L // 2 + 1).sum()) // 2)
        want.append([inter, dimers, len(res["sheets"]), res["largest_sheet"], res["peptides_in_sheets"], intra, 0, 0])
    got = d.sheet_observables()
    assert np.array_equal(got, np.array(want, dtype=np.int32)), (got, want)
    assert want[0][2] == 2 and want[0][3] == 4 and want[0][4] == 6 and want[0][5] == 1
    d.close()


def test_device_sheet_observables_hosttrace(tab, hosttrace_lib):
    check_device_sheet_observables(tab, hosttrace_lib)
