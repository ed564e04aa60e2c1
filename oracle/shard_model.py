"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-Python model of the SHARDED build's host logic (ropebwt2_b200/csrc/rb2_shard.inl), small
inputs only: the 36 sub-buckets (x, y) = BWT symbols whose suffix starts with "xy", their static
owners, the per-column table all-gather, whole-index ranks from local counts + the gathered totals
(the cross-bucket offsets of reference mrope.c:332-340), and the routing rule "a string in (x, y)
that inserts symbol a continues in (a, x)" with the target order (sub-bucket, y, source order).

It follows the batch restatement of mr_insert_multi_aux (reference mrope.c:184-233, in pre-column
coordinates; see oracle/bcr_oracle.c): per group, symbols are inserted in $,A,C,G,T,N (RCLO:
$,T,G,C,A,N, mrope.c:209-210) order at gL + sizes of the earlier symbols, and the next interval
start of the strings that inserted a is C_post[a] + occ_pre(a, insertion point).

Direct delivery is modelled too: from the gathered tables ALONE every rank derives,
for each of its output groups, the rank and the slot in that rank's next-column array where the group
belongs (the engine's PeerRoute: piece (a, x, y) goes from owner(x, y) to owner(a, x), the pieces of one
target in y order) and "stores" it there; the arrays assembled from those stores must equal the ones the
send/recv formulation (sort by (sub-bucket, y, source rank, source order)) produces.

Ranks talk through `comm.allgather(obj) -> [obj of rank 0, obj of rank 1, ...]` only, so the same
code runs in one process (world 1) and under a gloo process group (tests/test_dist_gloo.py).
"""
from __future__ import annotations


def owner_map(world: int):
    """Owner of sub-bucket s = x*6+y: contiguous ranges, the 16 ACGT x ACGT ones spread evenly
    (mirrors shard_owner_map in rb2_shard.inl; the CPU tests compare the two)."""
    own, k = [], 0
    for s in range(36):
        x, y = divmod(s, 6)
        own.append(min(world - 1, k * world // 16))
        if 1 <= x <= 4 and 1 <= y <= 4:
            k += 1
    return own


class LocalComm:
    """world size 1"""
    rank, world = 0, 1

    def allgather(self, obj):
        return [obj]


class ShardModel:
    def __init__(self, comm, so: int = 0):
        self.comm, self.so = comm, so
        self.rank, self.world = comm.rank, comm.world
        self.own = owner_map(self.world)
        self.seg = {s: [] for s in range(36) if self.own[s] == self.rank}   # my BWT segments
        self.tot = [[0] * 6 for _ in range(36)]                             # whole-index symbol totals, on every rank

    # ---- whole-index coordinates -------------------------------------------------------------
    def _start(self, s):          # global position of the first symbol of sub-bucket s
        return sum(sum(self.tot[t]) for t in range(s))

    def _occ(self, s, a, x_global):   # #a in BWT[0, x) for a position inside (or at the ends of) my sub-bucket s
        before = sum(self.tot[t][a] for t in range(s))
        return before + self.seg[s][:x_global - self._start(s)].count(a)

    # ---- one batch: every rank passes ITS strings (lists of nt6 codes, already reversed) ---------
    def insert_multi(self, my_strings):
        shares = self.comm.allgather([list(x) for x in my_strings])
        strings = [s + [0] for share in shares for s in share]              # global ids in (rank, position) order
        m = len(strings)
        if m == 0:
            return
        n0 = sum(self.tot[0])
        sorted_mode = self.so != 0
        order = [0, 4, 3, 2, 1, 5] if self.so == 2 else [0, 1, 2, 3, 4, 5]
        groups = []                                                          # my groups: [s, gL, gSize, members], in (s, position) order
        if self.own[0] == self.rank:
            groups = [[0, 0, n0, list(range(m))]] if sorted_mode else [[0, n0, 0, [k]] for k in range(m)]
        live, col = m, 0
        while live:
            # -- local: histograms, records, next groups (in the source order (a, s, local)) --------
            records = {}                                                     # s -> [(P, a, count)]
            nxt = {a: [] for a in range(1, 6)}                               # a -> [(s, P_of_record, gSize', members)]
            mem_tab = [[0] * 6 for _ in range(36)]
            grp_tab = [[0] * 6 for _ in range(36)]                            # groups I hand on, per (source sub-bucket, symbol)
            for s, gL, gSize, members in groups:
                by = {a: [k for k in members if strings[k][col] == a] for a in range(6)}
                P = gL
                for a in order:
                    sza = self._occ(s, a, gL + gSize) - self._occ(s, a, gL) if gSize else 0
                    if by[a]:
                        records.setdefault(s, []).append((P, a, len(by[a])))
                        mem_tab[s][a] += len(by[a])
                        if a:
                            nxt[a].append((s, P, sza, by[a]))
                            grp_tab[s][a] += 1
                    P += sza
            # -- gather the tables; post-column totals and bucket starts -------------------------------
            tabs = self.comm.allgather(mem_tab)
            gtabs = self.comm.allgather(grp_tab)
            post = [[self.tot[s][a] + sum(t[s][a] for t in tabs) for a in range(6)] for s in range(36)]
            cpost = [sum(sum(post[t]) for t in range(a * 6)) for a in range(6)]
            # -- ranks against the PRE-column index, then the merge -------------------------------------
            out = []                                                         # (target sub-bucket, y, gL', gSize', members)
            for a in range(1, 6):
                for s, P, sza, members in nxt[a]:
                    out.append(((a * 6 + s // 6), s % 6, cpost[a] + self._occ(s, a, P), sza, members))
            for s, recs in records.items():
                old, new, idx, base = self.seg[s], [], 0, self._start(s)
                for P, a, cnt in recs:                                       # sorted by position; new symbols go in front of old[P]
                    new += old[idx:P - base]
                    idx = P - base
                    new += [a] * cnt
                self.seg[s] = new + old[idx:]
            self.tot = post
            # -- exchange: every group goes to the owner of its next sub-bucket -----------------------------
            everything = self.comm.allgather(out)                            # (the engine sends each piece to its target only)
            mine = [(t, y, src, i, g) for src, lst in enumerate(everything) for i, (t, y, *g) in enumerate(lst) if self.own[t] == self.rank]
            mine.sort(key=lambda e: (e[0], e[1], e[2], e[3]))                # (sub-bucket, y, source rank, source order)
            groups = [[t, g[0], g[1], g[2]] for t, y, src, i, g in mine]
            # -- the same by direct delivery: (rank, slot) of every output group from the gathered tables -----
            route, cur = {}, [0] * self.world                                # (a, source sub-bucket) -> [target rank, next slot]
            for t in range(6, 36):
                a, x = divmod(t, 6)
                for y in range(6):
                    sb = x * 6 + y
                    ng = gtabs[self.own[sb]][sb][a]
                    if ng and self.own[sb] == self.rank:
                        route[(a, sb)] = [self.own[t], cur[self.own[t]]]
                    cur[self.own[t]] += ng
            stores = []                                                      # what my "merge epilogue" writes into the peers' arrays
            for (t, y, *g) in out:
                r = route[(t // 6, (t % 6) * 6 + y)]
                stores.append((r[0], r[1], [t, g[0], g[1], g[2]]))
                r[1] += 1
            landed = [(slot, g) for lst in self.comm.allgather(stores) for (dst, slot, g) in lst if dst == self.rank]
            landed.sort(key=lambda e: e[0])
            assert [slot for slot, _ in landed] == list(range(cur[self.rank])), "direct delivery: a slot filled twice or not at all"
            assert [g for _, g in landed] == groups, "direct delivery differs from the send/recv order"
            live = sum(len(g[2]) for lst in everything for (_, _, *g) in lst)
            col += 1

    # ---- the whole BWT (on every rank) -------------------------------------------------------------
    def text(self):
        parts = self.comm.allgather(self.seg)
        out = []
        for s in range(36):
            out += parts[self.own[s]][s]
        return out
