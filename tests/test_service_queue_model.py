"""Model check of the list-rebuild request queue (csrc/dmd_types.h SVC_Q_*, svc_request in dmd_engine.h,
svc_serve_in_kernel in dmd_cuda.cu): the device protocol restated as atomic steps of requester and server state
machines, run under many random interleavings.  Properties: every request ends served (word back to 0) or taken back
by its own warp and rebuilt in place -- never both, never twice; a server only works on a replica whose word it moved
1 -> 2; tickets are handed out in order; nothing is left pending and nobody waits forever.  No GPU needed: this checks
the PROTOCOL, the GPU tests check the code."""
import random

IDLE, ASKED, SERVING, TAKEN_BACK = 0, 1, 2, 3


class World:
    def __init__(self, n_rep, cap):
        self.flag = [IDLE] * n_rep
        self.head = self.tail = 0
        self.cap = cap
        self.ring = [0] * cap  # (ticket + 1) << 24 | replica, 0 = never written
        self.served = [0] * n_rep      # rebuilds done by a server, per replica
        self.in_place = [0] * n_rep    # rebuilds done by the warp itself after taking the request back
        self.busy = [None] * n_rep     # who is rebuilding the replica right now
        self.skipped = 0               # tickets whose slot had been overwritten when a server claimed them


def requester(w, rid, n_requests, patience, rng):
    """one replica's warp: n_requests list rebuilds, one after the other"""
    for _ in range(n_requests):
        w.flag[rid] = ASKED                      # st.release word = 1
        yield
        tk = w.tail                              # atomicAdd(tail, 1)
        w.tail += 1
        yield
        w.ring[tk % w.cap] = ((tk + 1) << 24) | rid  # st.release slot
        waited = 0
        while True:
            yield
            if w.flag[rid] == IDLE:              # served
                break
            waited += 1
            if w.flag[rid] == ASKED and waited > patience:  # cas 1 -> 3: nobody came, rebuild in place
                w.flag[rid] = TAKEN_BACK
                assert w.busy[rid] is None
                w.busy[rid] = "self"
                yield
                w.busy[rid] = None
                w.in_place[rid] += 1
                w.flag[rid] = IDLE
                break
        for _ in range(rng.randrange(3)):        # events until the next rebuild
            yield


def server(w, sid, done):
    """one service group"""
    while True:
        hd, tl = w.head, w.tail                  # relaxed loads
        yield
        if hd < tl:
            if w.head != hd:                     # atomicCAS(head, hd, hd + 1) lost
                continue
            w.head = hd + 1
            yield
            while (w.ring[hd % w.cap] >> 24) < hd + 1:   # the slot is written right after the ticket was taken
                yield
            if (w.ring[hd % w.cap] >> 24) != hd + 1:     # overwritten by a later lap of the ring: skip the ticket
                w.skipped += 1
                continue
            rid = w.ring[hd % w.cap] & 0xFFFFFF
            if w.flag[rid] != ASKED:             # cas 1 -> 2 failed: taken back (or already served through an older ticket)
                continue
            w.flag[rid] = SERVING
            assert w.busy[rid] is None
            w.busy[rid] = sid
            yield                                # ... the rebuild ...
            yield
            w.busy[rid] = None
            w.served[rid] += 1
            w.flag[rid] = IDLE                   # st.release word = 0
        elif done():
            return
        yield


def run(seed, n_rep, n_srv, n_requests, patience, cap=None):
    rng = random.Random(seed)
    w = World(n_rep, cap=cap or 2 * n_rep + 4)
    reqs = [requester(w, r, n_requests, patience, rng) for r in range(n_rep)]
    alive = set(range(n_rep))
    srvs = [server(w, "s%d" % s, lambda: not alive) for s in range(n_srv)]
    live_srv = set(range(n_srv))
    steps = 0
    while alive or live_srv:
        steps += 1
        assert steps < 200000, "no progress: somebody waits forever"
        pool = [("r", r) for r in alive] + [("s", s) for s in live_srv]
        kind, k = rng.choice(pool)
        try:
            next(reqs[k] if kind == "r" else srvs[k])
        except StopIteration:
            (alive if kind == "r" else live_srv).discard(k)
    return w


def check(w, n_requests):
    assert all(f == IDLE for f in w.flag)
    assert all(b is None for b in w.busy)
    for s, p in zip(w.served, w.in_place):
        assert s + p == n_requests            # every request exactly once, by a server or in place
    assert w.head == w.tail == len(w.flag) * n_requests  # every ticket claimed, none left for the next launch


def test_every_request_is_served_exactly_once():
    for seed in range(300):
        check(run(seed, n_rep=5, n_srv=2, n_requests=6, patience=10 ** 9), 6)


def test_requests_taken_back_leave_harmless_tickets():
    taken = 0
    for seed in range(300):
        w = run(seed, n_rep=6, n_srv=1, n_requests=5, patience=12)  # one slow server: some warps lose patience
        check(w, 5)
        taken += sum(w.in_place)
    assert taken > 0  # the scenario was exercised: stale tickets were claimed and skipped


def test_ring_overflow_costs_in_place_rebuilds_not_a_hang():
    """a service far too slow for the load: the tickets of requests taken back pile up until the ring laps itself"""
    skipped = 0
    for seed in range(200):
        w = run(seed, n_rep=6, n_srv=1, n_requests=8, patience=6, cap=7)
        check(w, 8)
        skipped += w.skipped
    assert skipped > 0  # the overflow was exercised


def test_no_server_at_all_degrades_to_in_place_rebuilds():
    w = run(1, n_rep=4, n_srv=0, n_requests=3, patience=5)
    assert w.served == [0] * 4 and w.in_place == [3] * 4
