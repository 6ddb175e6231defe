"""CPU check of the multi-GPU flag protocol of capital_b200/csrc/dist.cu (no GPU, no mirror implementation).

`capital_dist_trace_cholinv` dry-runs the REAL schedule code of one rank: every flag wait / signal, fused product, event
record / wait, peer DMA and arena window access it would enqueue for two consecutive cholinv::factor calls.  The traces of all
ranks of a grid are replayed here under CUDA's ordering rules (streams are FIFO; an event wait waits for the record that preceded
it; a flag wait stalls the stream until the flag reaches the value) and checked for
  * deadlock freedom: the replay must drain every stream of every rank;
  * data races: two accesses to overlapping windows of the same rank's arena, at least one a write, must be ordered by
    happens-before (vector clocks) -- i.e. every operand a product reads from a mirror slot was pushed AND waited for, and
    nothing of the next factorization overwrites what the previous one may still be reading.
"""
import ctypes as C
import itertools
import numpy as np
import pytest
import capital_b200 as cb
from capital_b200 import _lib

T_WAIT, T_SIGNAL, T_PRODUCT, T_EVREC, T_EVWAIT, T_DMA, T_KERNEL, T_READ, T_WRITE, T_MAT = range(1, 11)
NSTREAM = 12


def trace(size, rank, c, n, ci, bcm, split=1):
    g = cb.topo.square(size, rank, c).grid
    args = _lib.CholinvArgs(ci, split, bcm, b"U")
    cnt = C.c_int64()
    L = _lib.lib()
    st = L.capital_dist_trace_cholinv(C.byref(g), n, C.byref(args), None, 0, C.byref(cnt))
    assert st == 0
    buf = np.zeros((cnt.value, 8), dtype=np.int64)
    st = L.capital_dist_trace_cholinv(C.byref(g), n, C.byref(args), buf.ctypes.data_as(C.POINTER(C.c_int64)), cnt.value, C.byref(cnt))
    assert st == 0
    return buf


class Replay:
    def __init__(self, traces):
        self.P = len(traces)
        self.nclk = self.P * NSTREAM
        self.streams = [[[] for _ in range(NSTREAM)] for _ in range(self.P)]
        self.mats = [[] for _ in range(self.P)]
        for r, tr in enumerate(traces):
            for rec in tr:
                if rec[0] == T_MAT:
                    self.mats[r].append((int(rec[2]), int(rec[3]), int(rec[4])))
                else:
                    self.streams[r][int(rec[1])].append(tuple(int(v) for v in rec))
        self.head = [[0] * NSTREAM for _ in range(self.P)]
        self.clock = [[np.zeros(self.nclk, dtype=np.int64) for _ in range(NSTREAM)] for _ in range(self.P)]
        self.flags = {}      # (rank, word) -> (value, clock)
        self.events = {}     # (rank, ev) -> clock
        self.started = {}    # (rank, q, seq) -> clock at kernel start
        self.accesses = []   # (rank_of_memory, off, ld, rows, cols, is_write, group, clock, who)
        self.pending_acc = [[[] for _ in range(NSTREAM)] for _ in range(self.P)]

    def tick(self, r, s):
        c = self.clock[r][s].copy()
        c[r * NSTREAM + s] += 1
        self.clock[r][s] = c
        return c

    def partners(self, r, grid):
        c, d = grid
        z = r % c
        return [r - z + l for l in range(c) if l != z]

    def try_op(self, r, s, grid):
        q = self.streams[r][s]
        i = self.head[r][s]
        if i >= len(q):
            return False
        op = q[i]
        kind = op[0]
        clk = self.clock[r][s]
        if kind == T_WAIT:
            f = self.flags.get((r, op[2]))
            if f is None or f[0] < op[3]:
                return False
            self.clock[r][s] = np.maximum(clk, f[1])
            self.tick(r, s)
        elif kind == T_SIGNAL:
            c = self.tick(r, s)
            key = (op[2], op[3])
            old = self.flags.get(key)
            assert old is None or old[0] < op[4], f"flag {key} goes backwards: {old[0]} -> {op[4]}"
            self.flags[key] = (op[4], c if old is None else np.maximum(c, old[1]))
        elif kind == T_EVREC:
            c = self.tick(r, s)
            self.events[(r, op[2])] = c
        elif kind == T_EVWAIT:
            e = self.events.get((r, op[2]))
            if e is None:
                # CUDA: waiting for an event that was never recorded is a no-op -- but the schedule never does that on purpose
                own = [o for st in self.streams[r] for o in st if o[0] == T_EVREC and o[2] == op[2]]
                assert own, f"rank {r}: wait for event {op[2]} that is never recorded"
                return False
            self.clock[r][s] = np.maximum(clk, e)
            self.tick(r, s)
        elif kind == T_PRODUCT:
            # a GEMM launch: it needs nothing from other ranks while it runs (partials / final tiles are stored into the partners'
            # buffers, the handshake that follows is explicit flag traffic)
            self.tick(r, s)
            self.flush_accesses(r, s)
        elif kind in (T_READ, T_WRITE):
            self.pending_acc[r][s].append(op)
        else:  # T_DMA, T_KERNEL
            self.tick(r, s)
            self.flush_accesses(r, s)
        self.head[r][s] = i + 1
        return True

    def flush_accesses(self, r, s, start=None):
        end = self.clock[r][s]
        start = end if start is None else start
        for op in self.pending_acc[r][s]:
            self.accesses.append((op[2], op[3], op[4], op[5], op[6], op[0] == T_WRITE, op[7] if op[0] == T_WRITE else 0, start, (r, s), end))
        self.pending_acc[r][s] = []

    def run(self, grid):
        progress = True
        while progress:
            progress = False
            for r in range(self.P):
                for s in range(NSTREAM):
                    while self.try_op(r, s, grid):
                        progress = True
        stuck = [(r, s, self.streams[r][s][self.head[r][s]]) for r in range(self.P) for s in range(NSTREAM)
                 if self.head[r][s] < len(self.streams[r][s])]
        return stuck

    # ---- race detection ----
    def slot_of(self, rank, off):
        for base, ld, cols in self.mats[rank]:
            if base <= off < base + ld * cols * 8:
                return base, ld
        return None

    def races(self):
        buckets = {}
        for a in self.accesses:
            rank, off, ld, rows, cols = a[:5]
            sl = self.slot_of(rank, off)
            if sl is None or sl[1] != ld:
                key, r0, c0 = (rank, off, "raw"), 0, 0   # 1-D buffers (gather slots): whole-buffer granularity per base address
                rect = (off, off + rows * cols * 8, 0, 1)
                key = (rank, "raw")
            else:
                e = (off - sl[0]) // 8
                c0, r0 = divmod(e, ld)
                rect = (r0, r0 + rows, c0, c0 + cols)
                key = (rank, sl[0])
            buckets.setdefault(key, []).append((rect, a))
        bad = []
        for key, lst in buckets.items():
            for (ra, a), (rb, b) in itertools.combinations(lst, 2):
                if not (a[5] or b[5]):
                    continue
                if a[5] and b[5] and a[6] and a[6] == b[6]:
                    continue  # cooperating writers of one fused product (disjoint tiles)
                if ra[0] >= rb[1] or rb[0] >= ra[1] or ra[2] >= rb[3] or rb[2] >= ra[3]:
                    continue
                wa, wb = a[8], b[8]
                ia, ib = wa[0] * NSTREAM + wa[1], wb[0] * NSTREAM + wb[1]
                a_before_b = a[9][ia] <= b[7][ia]   # a's end happens-before b's start
                b_before_a = b[9][ib] <= a[7][ib]
                if not (a_before_b or b_before_a):
                    bad.append((key, ra, "W" if a[5] else "R", wa, rb, "W" if b[5] else "R", wb))
        return bad


GRIDS = {2: (2, 1), 4: (1, 2), 8: (2, 2)}


@pytest.mark.parametrize("size,n,ci,bcm", [
    (2, 1024, 0, -3), (2, 1024, 1, -2),
    (4, 1024, 0, -2), (4, 2048, 1, -3),
    (8, 1024, 1, -2), (8, 2048, 0, -3), (8, 4096, 1, -3),
])
def test_flag_protocol_is_deadlock_free_and_race_free(size, n, ci, bcm, monkeypatch):
    # small nodes must exercise the deferred class too
    monkeypatch.setenv("CAPITAL_DIST_FAR_MIN", "64")
    monkeypatch.setenv("CAPITAL_DIST_SIDE_MIN", "32")
    monkeypatch.setenv("CAPITAL_DIST_CHUNK_MIN", "256")  # ... and the chunked, pushed products
    c, d = GRIDS[size]
    traces = [trace(size, r, c, n, ci, bcm) for r in range(size)]
    rp = Replay(traces)
    stuck = rp.run((c, d))
    assert not stuck, f"deadlock: {len(stuck)} streams blocked, e.g. {stuck[:4]}"
    kinds = np.concatenate(traces)[:, 0]
    if size > 1:
        assert (kinds == T_PRODUCT).sum() > 0
    if d > 1:
        assert (kinds == T_DMA).sum() > 0 and (kinds == T_WAIT).sum() > 0
    bad = rp.races()
    assert not bad, f"{len(bad)} unordered conflicting accesses, e.g. {bad[:3]}"


@pytest.mark.parametrize("size,n,ci", [(8, 2048, 1), (8, 2048, 0), (4, 2048, 1), (2, 1024, 0)])
def test_uneven_split_is_clean_too(size, n, ci, monkeypatch):
    """split = 2 (cholinv.hpp:92,107: the left child gets a quarter): node sizes stop being powers of two, every window offset changes"""
    monkeypatch.setenv("CAPITAL_DIST_FAR_MIN", "64")
    monkeypatch.setenv("CAPITAL_DIST_SIDE_MIN", "32")
    monkeypatch.setenv("CAPITAL_DIST_CHUNK_MIN", "256")
    c, d = GRIDS[size]
    traces = [trace(size, r, c, n, ci, -3, split=2) for r in range(size)]
    rp = Replay(traces)
    assert not rp.run((c, d))
    assert not rp.races()


def test_pipelined_chunked_products_are_clean_too(monkeypatch):
    """opt-in CAPITAL_DIST_PIPELINE=1: chunk j + 1 is issued before chunk j is added up (three rotating exchange sets + the
    'reduced' flag keep a partner from overwriting a set that is still being read)"""
    monkeypatch.setenv("CAPITAL_DIST_PIPELINE", "1")
    monkeypatch.setenv("CAPITAL_DIST_CHUNK_MIN", "256")
    monkeypatch.setenv("CAPITAL_DIST_FAR_MIN", "64")
    monkeypatch.setenv("CAPITAL_DIST_SIDE_MIN", "32")
    traces = [trace(8, r, 2, 2048, 1, -3) for r in range(8)]
    rp = Replay(traces)
    assert not rp.run((2, 2))
    assert not rp.races()
    # and the flag that protects the rotating sets is load-bearing: without its waits the replay finds the race
    CTRL_RED_LO, CTRL_RED_HI = 192, 272
    mut = [tr[~((tr[:, 0] == T_WAIT) & (tr[:, 2] >= CTRL_RED_LO) & (tr[:, 2] < CTRL_RED_HI))] for tr in traces]
    rp2 = Replay(mut)
    assert not rp2.run((2, 2))
    assert rp2.races()


def test_single_stream_schedule_also_clean(monkeypatch):
    monkeypatch.setenv("CAPITAL_DIST_TWO_STREAM", "0")
    traces = [trace(8, r, 2, 1024, 1, -2) for r in range(8)]
    rp = Replay(traces)
    assert not rp.run((2, 2))
    assert not rp.races()
