#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native PRIME20 DMD engine.

Metric (BASELINE.json): DMD collision events/s per B200 (every processed calendar event, counted like the
reference's `coll`, main.F90:639).  Workload at N=1: BASELINE config 2 -- the 48-peptide Abeta16-22 (KLVFFAE)
PRIME20 box, N = 1344 beads, L = 158.54 A, T* = 0.18, Andersen thermostat on (-Dcanon) -- as an ensemble of R
independent replicas resident on one GPU (four replicas per hardware warp, 8 lanes each, in lockstep; the replica
count is the one that fills the device in one wave next to the list-rebuild service CTAs of the same kernel,
dmdb_device_fill).  N>1: BASELINE config 3 -- the same box count per GPU (weak scaling), but the replicas form
11-temperature ladders (temp_018 ... temp_050, qfile/script.sh:11-18) whose members are striped across the GPUs, with
one replica-exchange step per bench step: dmdb_exchange = energy reduction + ncclAllGather of (E_pot, T*) over NVLink
+ Metropolis decision + temperature change of the swapped replicas, all on the device and inside the timed region.

A "step" = every replica advances `--events` calendar events (one launch of the persistent event-loop kernel).

  value : whole-job events/s with state resident in HBM (device time of the event-loop kernel, CUDA events on the
          library's stream, max over ranks)
  e2e   : the same metric through the C ABI with HOST buffers inside the timed region: dmdb_set_state_all (H2D of
          every replica's restart state from pinned memory, run start: nbor + events) -> dmdb_run ->
          dmdb_sync_positions -> dmdb_get_state_all + dmdb_potential_energies (D2H)
  --impl reference : the reference's CPU implementation of the path.  The Fortran cannot be built in this image
          (no Fortran compiler, Intel IFPORT, MPI), so this times the C++ oracle port of it on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DMD collision events/sec per B200"
UNIT = "events/s"
WORKLOAD = "config2: 48 x Abeta16-22 (KLVFFAE) PRIME20 box, N=1344 beads, L=158.54 A, T*=0.18, canon"
WORKLOAD_LADDER = ("config3: 48 x Abeta16-22 (KLVFFAE) PRIME20 boxes, N=1344 beads, L=158.54 A, canon, as 11-temperature "
                   "ladders T*=0.18..0.50 (temp_018..temp_050) striped across the GPUs, replica exchange every step")
TSTAR, BOXL = 0.18, 158.54


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (0 = dmdb_device_fill: one 28-warp CTA (112 replicas) per "
                    "SM minus the list-rebuild service CTAs, 14112 on a 148-SM B200)")
    ap.add_argument("--events", type=int, default=20000, help="calendar events per replica per step")
    ap.add_argument("--ref-events", type=int, default=400000, help="events per host thread per step (--impl reference)")
    ap.add_argument("--ladder", action="store_true", help="config 3 on one GPU too: 11-temperature ladders + exchange per step "
                    "(always on for --gpus > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (single trajectory, "
                    "12 288-bead box, bulk kernels on a 10^6-bead box)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha-256 (first 16 hex digits) over the CUDA sources of the library with comments and white space removed: ties a
    measured DRAM-traffic figure to the CODE it was captured from (the GPU box holds no .git); editing a comment does
    not invalidate a capture, editing a statement does"""
    import hashlib
    import re
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if not name.endswith((".h", ".cu")):
            continue
        with open(os.path.join(csrc, name), "r", encoding="utf-8", errors="replace") as f:
            text = f.read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)  # block comments
        text = re.sub(r"//[^\n]*", " ", text)               # line comments (the sources hold no '//' inside string literals)
        text = re.sub(r"\s+", " ", text)
        h.update(name.encode() + b"\0" + text.encode())
    return h.hexdigest()[:16]


def algorithmic_bytes(n_beads, d_events, d_pair, d_ghost, d_visits):
    """a committed pair event moves 2*64 (records) + 2*48 + 2*16 (results) + 84 B per neighbour-list
    entry visited (4 B entry + 64 B partner record + 16 B calendar entry); an interval pseudo-event (advance_sync)
    moves 48N + 24N + 16(N+3).  List rebuilds are NOT counted (conservative)."""
    interval = max(d_events - d_pair - d_ghost, 0)
    return d_pair * 256 + d_visits * 84 + d_ghost * (64 + 84 * 10) + interval * (72 * n_beads + 16 * (n_beads + 3))


def cpu_baseline_sample(tab, topo, sv, seconds_target=12.0):
    """the oracle (port of the reference's serial algorithm) on ONE host core, bounded sample of the same workload.
    Both builds of the port are timed -- the parity build (g++ -O2 -ffp-contract=off) and the speed build (-O3
    -march=x86-64-v3, FMA allowed; oracle/Makefile) -- and the FASTER one is the baseline."""
    from oracle.binding import OracleDMD
    from parallel_dmd_for_biomolecules_b200 import tables
    rates = {}
    for name, fast in (("parity build -O2 -ffp-contract=off", False), ("speed build -O3 -march=x86-64-v3", True)):
        o = OracleDMD(tables.make_params(boxl=BOXL, tstar=TSTAR, canon=True), topo, tab, fast=fast)
        o.set_state(sv)
        o.run(100000)
        n, sec = 0, 0.0
        while sec < seconds_target / 2:
            sec += o.run(500000)
            n += 500000
        rates[name] = (n / sec, n, sec)
        o.close()
    best = max(rates, key=lambda k: rates[k][0])
    return {"value": rates[best][0], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d events of one replica of the workload after 1e5 warm-up events, C++ oracle, %s (%.1f s); the other "
                      "build: %s" % (rates[best][1], best, rates[best][2],
                                     ", ".join("%s %.3g events/s" % (k, v[0]) for k, v in rates.items() if k != best))}


def fastest_oracle_build(tab, topo, sv):
    """which build of the port is faster on this host (a 2 x 2e5-event probe on one core)"""
    from oracle.binding import OracleDMD
    from parallel_dmd_for_biomolecules_b200 import tables
    r = {}
    for fast in (False, True):
        o = OracleDMD(tables.make_params(boxl=BOXL, tstar=TSTAR, canon=True), topo, tab, fast=fast)
        o.set_state(sv)
        o.run(50000)
        r[fast] = 200000 / o.run(200000)
        o.close()
    return r[True] > r[False], r


def extras(tab, topo_b, sv_b, peak):
    """Secondary measurements reported beside the headline (rank 0, 1 GPU): the same event loop on ONE trajectory
    with the CTA-per-replica engine (batched conservative commit), BASELINE config 4 (192 chains x 16 residues,
    12 288 beads) and the bulk kernels on BASELINE config 5 (~10^6 beads) with their algorithmic HBM traffic
    (events() moves 64N + 68 P_up + 16N bytes; nbor() 24N + 4N + 8 P_up plus the candidate scan)."""
    from parallel_dmd_for_biomolecules_b200 import genconfig, tables
    from parallel_dmd_for_biomolecules_b200.dmd import DMD
    out = {}
    for name, eng in (("warp_per_replica", 1), ("cta_per_replica_batched_commit", 2)):
        d = DMD(tables.make_params(boxl=BOXL, tstar=TSTAR, canon=True, n_replicas=1, engine=eng), topo_b, tab)
        d.set_state(sv_b)
        d.run(20000)
        n = 100000 if eng == 2 else 30000
        st = d.run(n)
        out["single_trajectory_config2_" + name] = {"events_per_s": n / (st.device_ms * 1e-3)}
        if eng == 2:
            bs = d.batch_stats()
            out["single_trajectory_config2_" + name]["events_committed_per_round"] = (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1)
        d.close()
    topo4, sv4 = genconfig.generate_box(["KLVFFAEKLVFFAEKL"], [192], 200.0, 0.3, tab, seed=3)
    for name, eng in (("cta_per_replica", 2), ("whole_gpu", 3)):
        d = DMD(tables.make_params(boxl=200.0, tstar=0.3, canon=True, n_replicas=1, engine=eng), topo4, tab)
        d.set_state(sv4)
        d.run(5000)
        t0 = time.perf_counter()
        d.run(100000)
        dt = time.perf_counter() - t0
        bs = d.batch_stats()
        out["config4_12288_beads_one_trajectory_" + name] = {
            "events_per_s": 100000 / dt, "events_committed_per_round": (bs["executed"] - bs["rolled_back"]) / max(bs["rounds"], 1)}
        d.close()
    nch = 35715
    boxl = BOXL * (nch / 48.0) ** (1.0 / 3.0)
    topo5, sv5 = genconfig.generate_box(["KLVFFAE"], [nch], boxl, 0.5, tab, seed=5)
    N5 = topo5.n_beads
    d = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, engine=1, nbr_capacity=32), topo5, tab)
    d.set_state(sv5)
    p_up = int(d.nbors(0)[0][-1])
    res = {"beads": N5, "up_pairs": p_up}
    for name, fn, nbytes in (("events", d.events, 64 * N5 + 68 * p_up + 16 * N5), ("nbor", d.nbor, 28 * N5 + 64 * N5 + 2 * 8 * p_up)):
        ms = []
        for _ in range(5):
            fn()
            ms.append(d.stats().device_ms)
        t = min(ms[1:])
        res[name] = {"device_ms": t, "algorithmic_GB_per_s": nbytes / (t * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (t * 1e-3) / 1e9 / peak}
    out["config5_bulk_kernels"] = res
    d.close()
    # the same box through the event loop: engine 3, every round of the batched commit spread over the whole GPU
    d = DMD(tables.make_params(boxl=boxl, tstar=0.5, canon=True, n_replicas=1, engine=3, nbr_capacity=32), topo5, tab)
    d.set_state(sv5)
    d.run(200000)
    b0 = d.batch_stats()
    t0 = time.perf_counter()
    d.run(2000000)
    dt = time.perf_counter() - t0
    b1 = d.batch_stats()
    out["config2_aggregated_regime"] = aggregated_regime(tab, peak)
    out["config5_1e6_beads_one_trajectory_whole_gpu_engine"] = {
        "events_per_s": 2000000 / dt, "timing": "wall clock around dmdb_run (kernel relaunches at pseudo-events included)",
        "events_committed_per_round": (b1["executed"] - b1["rolled_back"] - b0["executed"] + b0["rolled_back"]) / max(b1["rounds"] - b0["rounds"], 1)}
    d.close()
    return out


def aggregated_regime(tab, peak):
    """The same kernel in the regime the reference spends its 2e10 events in: every replica starts
    from a PRE-AGGREGATED 48-peptide box -- the state tools/make_aggregated_fixture.py annealed on the device through
    the reference's schedule (qfile/script.sh:11-14) and committed as tests/golden/aggregated_L80.npz -- with its own
    random-number stream, and runs 1e5 + 1e6 events at T* = 0.18.  Same roofline arithmetic as the headline."""
    from parallel_dmd_for_biomolecules_b200 import genconfig, tables
    from parallel_dmd_for_biomolecules_b200.dmd import DMD, device_fill
    path = os.path.join(ROOT, "tests", "golden", "aggregated_L80.npz")
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/aggregated_L80.npz missing"}
    fx = np.load(path)
    boxl = float(fx["boxl"])
    topo, _ = genconfig.system_b(tab, TSTAR, seed=1, boxl=boxl)
    # an aggregated box rebuilds its lists ~5 x slower than a dilute one (a bead has many more candidates): the measured
    # optimum is 44 list-rebuild service CTAs beside 104 event-loop CTAs (tools/aggr_run.py: 36: 1.26e8, 40: 1.38e8, 44: 1.48e8, 48: 1.45e8) (22 / 126 for the dilute headline)
    import torch
    service = 44
    fill_r, fill_s = device_fill(0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    R = (sms - service) * (fill_r // (sms - fill_s))  # replicas per event-loop CTA as dmdb_device_fill sizes them
    d = DMD(tables.make_params(boxl=boxl, tstar=TSTAR, canon=True, n_replicas=R, seed=77001), topo, tab)
    d.set_service_ctas(service)
    d.set_state(np.ascontiguousarray(fx["sv"]), np.ascontiguousarray(fx["bptnr"]))
    d.run(100000)
    s0 = d.stats()
    n_ev = 1000000
    st = d.run(n_ev)
    s1 = d.stats()
    so = d.sheet_observables()
    nb = (len(d.nbors(0, False)[1]) + len(d.nbors(0, True)[1])) / topo.n_beads
    d_events, d_pair = s1.events - s0.events, s1.pair_events - s0.pair_events
    hot = sum(s1.nevents[k] - s0.nevents[k] for k in (1, 2, 3))
    abytes = algorithmic_bytes(topo.n_beads, d_events, d_pair, s1.ghosts - s0.ghosts, s1.nbr_visits - s0.nbr_visits)
    gbs = abytes / (st.device_ms * 1e-3) / 1e9
    d.close()
    return {"events_per_s": R * n_ev / (st.device_ms * 1e-3), "replicas": R, "service_ctas": service, "events_per_replica": n_ev, "box_A": boxl,
            "start": "tests/golden/aggregated_L80.npz: 48 x KLVFFAE at 8 x the reference concentration, annealed on the device "
                     "(T* 0.50 -> 0.22: %d events each, then %d events at 0.18; tools/make_aggregated_fixture.py)" % (
                         int(fx["events_per_anneal_T"]), int(fx["events_at_018"])),
            "n_bar_neighbours_per_bead": nb, "list_entries_visited_per_event": (s1.nbr_visits - s0.nbr_visits) / max(d_events, 1),
            "inter_chain_hbonds_mean": float(so[:, 0].mean()), "largest_sheet_mean": float(so[:, 3].mean()),
            "peptides_in_sheets_mean": float(so[:, 4].mean()),
            "cold_path_fraction_of_pair_events": 1.0 - hot / max(d_pair, 1),
            "list_rebuilds_per_1e3_events": 1e3 * ((s1.updates + s1.forced_updates) - (s0.updates + s0.forced_updates)) / max(d_events, 1),
            "algorithmic_GB_per_s": gbs, "frac_of_measured_hbm": gbs / peak}


def run_reference(args, rank, world):
    """--impl reference: the oracle port on every host core (one independent replica per thread)."""
    if rank != 0:
        return
    from oracle.binding import OracleDMD
    from parallel_dmd_for_biomolecules_b200 import genconfig, tables
    tab = tables.load_default_tables()
    topo, sv = genconfig.system_b(tab, TSTAR, seed=1)
    T = os.cpu_count() or 1
    use_fast, probe = fastest_oracle_build(tab, topo, sv)
    reps = []
    for k in range(T):
        o = OracleDMD(tables.make_params(boxl=BOXL, tstar=TSTAR, canon=True, seed=1058472402 + k), topo, tab, fast=use_fast)
        o.set_state(sv)
        reps.append(o)

    def step():
        th = [threading.Thread(target=o.run, args=(args.ref_events,)) for o in reps]  # ctypes releases the GIL
        for t in th:
            t.start()
        for t in th:
            t.join()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = T * args.ref_events * args.steps / dt
    sample = ("%d host threads x %d events per step, one replica of the workload per thread; %s build of the port (1-core probe: "
              "parity %.3g, speed %.3g events/s)" % (T, args.ref_events, "speed (-O3 -march=x86-64-v3)" if use_fast else
                                                     "parity (-O2 -ffp-contract=off)", probe[False], probe[True]))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "the Fortran reference cannot be compiled in this image; this is its C++ "
                   "oracle port (oracle/), serial semantics, one trajectory per host thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist

    from parallel_dmd_for_biomolecules_b200 import genconfig, replica_exchange, tables
    from parallel_dmd_for_biomolecules_b200.dmd import DMD, device_fill

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tab = tables.load_default_tables()
    topo, sv = genconfig.system_b(tab, TSTAR, seed=1)
    fill_replicas, service_ctas = device_fill(local_rank)
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    per_cta = fill_replicas // (sms - service_ctas)  # replicas per 28-warp event-loop CTA as dmdb_device_fill sizes them
    per_warp = per_cta // 28
    if args.replicas <= 0:
        args.replicas = fill_replicas
    N, R, E = topo.n_beads, args.replicas, args.events
    p = tables.make_params(boxl=BOXL, tstar=TSTAR, canon=True, n_replicas=R, device=local_rank, seed=1058472402 + 100003 * rank)
    d = DMD(p, topo, tab)
    d.set_state(sv)
    ladder = world > 1 or args.ladder
    xch = {"ms": 0.0, "attempted": 0, "accepted": 0, "changed": 0, "ladders": 0, "launches": 0}
    if ladder:  # BASELINE config 3: the reference's 11 temperatures, ladders striped across the GPUs
        replica_exchange.init_communicator(d)  # the library's own NCCL communicator (dmdb_comm_init)
        d.apply_temperatures(replica_exchange.ladder_temperatures(world, rank, R))

    def step(k, timed=False):
        ms = d.run(E).device_ms
        if ladder:  # dmdb_exchange: energies -> ncclAllGather -> decision -> retemp, on the library's stream
            st = d.exchange(k, seed=4242, ladder_size=len(replica_exchange.LADDER))
            ms += st.device_ms
            if timed:
                xch["ms"] += st.device_ms
                xch["attempted"] += st.attempted
                xch["accepted"] += st.accepted
                xch["changed"] += st.changed_local
                xch["ladders"] = st.ladders
                xch["launches"] += st.kernel_launches
        return ms

    for k in range(args.warmup):
        step(k)
    s0 = d.stats()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for k in range(args.steps):
        dev_ms += step(args.warmup + k, timed=True)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    s1 = d.stats()
    # ---- e2e through the C ABI with host buffers (pinned), copies inside the timed region
    d.sync_positions()
    host_sv = torch.empty((R, N, 6), dtype=torch.float64, pin_memory=True).numpy()
    host_bp = torch.empty((R, N), dtype=torch.int32, pin_memory=True).numpy()
    d.get_state_all(host_sv, host_bp)

    def e2e_step():
        d.set_state_all(host_sv, host_bp)
        d.run(E)
        d.sync_positions()
        d.get_state_all(host_sv, host_bp)
        return d.potential_energies()[0]

    e2e_step()
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        epot = e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    h2d = R * (N * 48 + N * 4 + 8)  # raw sv + bptnr + temperatures; every derived array is built on the device
    d2h = R * (N * 48 + N * 4) + R * 64 + 2 * R * 512  # sv + bptnr, energy records, per-replica scalars (error words)
    # ---- reduce over ranks
    vals = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_max, e2e_wall_max = [float(x) for x in vals.cpu()]
    d_events = s1.events - s0.events
    d_pair = s1.pair_events - s0.pair_events
    d_ghost = s1.ghosts - s0.ghosts
    d_visits = s1.nbr_visits - s0.nbr_visits
    total_events = world * R * E * args.steps
    value = total_events / (dev_ms_max * 1e-3)
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes(N, d_events, d_pair, d_ghost, d_visits)
    achieved = abytes / (dev_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: from the tracked ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of
    # one `ncu --set full` launch, per event) -- used only while the kernel sources are the ones it was captured from
    traffic, traffic_note = None, "no capture"
    try:
        with open(os.path.join(ROOT, "profiles", "event_loop_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("source_sha16") == kernel_source_hash():
            traffic = tj["dram_bytes_per_event"] * R * E
            traffic_note = "%s; %.0f B/event x events of one launch" % (tj["source"], tj["dram_bytes_per_event"])
        else:
            traffic_note = "stale: profiles/event_loop_traffic.json was captured from other kernel sources (%s != %s)" % (
                tj.get("source_sha16"), kernel_source_hash())
    except Exception as e:  # noqa: BLE001
        traffic_note = "unavailable: %s" % e
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if not ladder else WORKLOAD_LADDER, "replicas_per_gpu": R, "beads_per_replica": N,
                       "events_per_replica_per_step": E,
                       "parallelism": "%d replicas per hardware warp (%d lanes each, in lockstep), %d replicas per GPU, %d GPU(s)%s; "
                                      "per GPU %d event-loop CTAs + %s list-rebuild service CTAs in ONE kernel" % (
                           per_warp, 32 // per_warp, R, world,
                           ", dmdb_exchange (ncclAllGather + device-side decision + retemp) per step" if ladder else "",
                           (R + per_cta - 1) // per_cta, service_ctas if R == fill_replicas else "auto"),
                       "exchange": None if not ladder else {
                           "ladders": xch["ladders"], "ladder_size": len(replica_exchange.LADDER),
                           "pairs_attempted_per_step": xch["attempted"] / args.steps,
                           "swap_acceptance": xch["accepted"] / max(xch["attempted"], 1),
                           "local_replicas_retempered_per_step": xch["changed"] / args.steps,
                           "device_ms_per_step": xch["ms"] / args.steps,
                           "what": "energy kernel + pack + ncclAllGather + decide + select + retemp kernels on the library's stream, inside ms_per_step"},
                       "l2": "no flush needed: resident working set per GPU %.1f GB >> 126 MB L2" % (R * N * 1100 / 1e9),
                       "event_count_convention": "all calendar events incl. ghost/interval pseudo-events (main.F90:639)",
                       "pair_event_fraction": d_pair / max(d_events, 1),
                       "mean_list_entries_visited_per_event": d_visits / max(d_events, 1),
                       "single_trajectory_events_per_s": value / (world * R),
                       "wall_ms_per_step": 1e3 * wall_max / args.steps},
            "clocks": clocks,
            "e2e": {"value": world * R * E * e2e_steps / e2e_wall_max, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "path": "dmdb_set_state_all(pinned host sv,bptnr) -> dmdb_run -> dmdb_sync_positions -> "
                            "dmdb_get_state_all + dmdb_potential_energies"},
            "gpu_launches": args.steps + xch["launches"],  # event-loop kernel per step (+ the exchange's own kernels)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src, "kernel": "dmd_event_loop_kernel",
                         "algorithmic_bytes_per_launch": abytes / args.steps,
                         "note": "bound by warp-instruction issue under dependent-gather and instruction-fetch stalls (IPC 0.4 per scheduler at 28 warps per SM), not by HBM: see DESIGN.md section 4"},
        }
        if world == 1 and not args.no_extras:
            d.close()
            out["extras"] = extras(tab, topo, sv, peak)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(tab, topo, sv)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
