"""CPU model of the look-ahead kernel's hazard protocol (simplexoncuda_b200/csrc/b2s_lookahead.cuh).

The streaming CTAs overwrite the tableau in place while the helper CTAs read rows 0 / 1+q' and column p' of it for the NEXT
pivot.  The protocol that keeps them apart has no locks: one atomic word [ticket | published row | published column | quiet],
publications that are or-ed into it by every helper (each keeping the ticket count its own atomic returned), "leave the row
alone / hold the column old" decided by the claimer from the value its claiming atomic returned, and a completion record per
tile that says what the claimer did.  This test replays that protocol as an event simulation under random interleavings --
exact rational arithmetic, so "same bits" is "same value" -- and checks, for every interleaving, that

  * the helpers' col' (row 1+q'), b' (row 0) and rowp' (column p') equal the sequentially updated tableau's,
  * the stored tableau equals the sequentially updated one everywhere except (possibly) the held-old column p',
  * after the write-back of rowp' (la_flush_kernel) it is equal everywhere.

It is a model of the device logic, not the device code: the CUDA kernel itself is compared with the oracle bit for bit in
tests/test_gpu_parity.py (stepping, chunked iterate, helper counts, variants) and tests/test_sharded_same_gpu.py."""
import random
from fractions import Fraction

import pytest

NONE_COL = -2     # "no column will be published" sentinel (kNoColumn)


class Word:
    """The 64-bit ticket word: atomics are single events, so every actor sees a total order of modifications."""

    def __init__(self):
        self.ticket = 0
        self.row = -1
        self.col = -1
        self.quiet = False

    def claim(self):
        old = (self.ticket, self.row, self.col, self.quiet)
        self.ticket += 1
        return old

    def or_row(self, r):
        old = self.ticket
        self.row = r
        return old

    def or_col(self, c):
        old = self.ticket
        self.col = c
        return old


def run_trial(seed):
    rng = random.Random(seed)
    R, m = rng.randint(9, 40), rng.choice([8, 16, 24])
    tile_rows, chunk_cols = rng.choice([2, 4]), 8
    nchunks = m // chunk_cols
    G, H = rng.randint(1, 6), rng.randint(1, 3)            # streaming CTAs with an implicit first tile, helpers
    base = G
    T = [[Fraction(rng.randint(-4, 4)) for _ in range(m)] for _ in range(R)]
    lp = rng.randrange(m)                                    # pivot column of the running update
    piv = Fraction(rng.choice([1, 2, 4]))
    s = [Fraction(rng.randint(-3, 3)) for _ in range(m)]
    s[lp] = Fraction(0)
    skip = rng.random() < 0.7                                # skip_zero_rows
    rowp = [Fraction(rng.choice([0, 0, 1, -2, 3]) if skip else rng.randint(1, 3)) for _ in range(R)]
    live = [r for r in range(1, R) if (not skip) or rowp[r] != 0]
    pos = {r: k for k, r in enumerate(live)}
    ntr = (len(live) + tile_rows - 1) // tile_rows
    ntiles = ntr * nchunks
    reverse = rng.random() < 0.5
    rq = rng.randrange(1, R)                                 # row of the next entering variable
    lpn = lp if rng.random() < 0.15 else rng.randrange(m)    # next pivot column (sometimes the same constraint again)
    same_col = lpn == lp

    # sequential truth
    new = [[(rowp[r] / piv) if i == lp else T[r][i] + s[i] * rowp[r] for i in range(m)] for r in range(R)]

    word = Word()
    rec = {}                                                 # tile -> (done, row_held, col_held)

    def tile_of(ticket):
        tmap = ntiles - 1 - ticket if reverse else ticket
        return tmap // nchunks, tmap % nchunks             # (row block, chunk)

    def ticket_of(rb, chunk):
        tmap = rb * nchunks + chunk
        return ntiles - 1 - tmap if reverse else tmap

    # ---- streaming CTAs -------------------------------------------------------------------------------------------
    class Cta:
        def __init__(self, g):
            self.cur = (g, -1, -1, True) if g < ntiles else None   # implicit tile: ticket g, saw nothing, wants a record
            self.nxt = None
            self.state = "claim"

        def step(self):
            if self.state == "claim":                        # the prefetching claim at the top of a tile
                t, row, col, quiet = word.claim()
                tk = t + base
                self.nxt = (tk, row, col if col != NONE_COL else -1, (not quiet) and col != NONE_COL) if tk < ntiles else None
                self.state = "work" if self.cur else "advance"
            elif self.state == "work":                       # the tile's loads, FMAs and stores
                tk, row, col, want = self.cur
                rb, chunk = tile_of(tk)
                for k in range(rb * tile_rows, min((rb + 1) * tile_rows, len(live))):
                    r = live[k]
                    if r == row:
                        continue                             # left to the helpers
                    for i in range(chunk * chunk_cols, (chunk + 1) * chunk_cols):
                        if i == lp:
                            T[r][i] = rowp[r] / piv          # pivot column: a_pr / pivot
                        elif i != col:                       # held-old lane keeps its value
                            T[r][i] = T[r][i] + s[i] * rowp[r]
                self.pending = (tk, row >= 0, col >= 0) if want else None
                self.state = "record"
            elif self.state == "record":                     # the record goes out during the next tile
                if self.pending:
                    rec[self.pending[0]] = (True, self.pending[1], self.pending[2])
                self.state = "advance"
            elif self.state == "advance":
                self.cur, self.nxt = self.nxt, None
                self.state = "claim" if self.cur else "done"
            return self.state != "done"

    # ---- helpers --------------------------------------------------------------------------------------------------
    blocks = list(range(nchunks))                            # one ratio block per chunk in this model
    out = {"b": [None] * m, "col": [None] * m, "rowp": [None] * R}

    class Helper:
        def __init__(self, h):
            self.h = h
            self.todo = []
            # stage 0: row 0 (not in the list)
            for b in blocks[h::H]:
                self.todo.append(("row0", b))
            self.todo.append(("pub_row",))
            for b in blocks[h::H]:
                self.todo.append(("ratio", b))
            self.todo.append(("barrier_R",))
            self.todo.append(("pub_col",))
            for r in range(h, R, H):
                self.todo.append(("gather", r))
            self.todo.append(("barrier_G",))
            if h == 0:
                self.todo.append(("quiet",))
            self.c_row = self.c_col = None

        def step(self):
            if not self.todo:
                return False
            op = self.todo[0]
            if op[0] == "row0":
                for i in range(op[1] * chunk_cols, (op[1] + 1) * chunk_cols):
                    T[0][i] = rowp[0] / piv if i == lp else T[0][i] + s[i] * rowp[0]
                    out["b"][i] = T[0][i]
            elif op[0] == "pub_row":
                self.c_row = word.or_row(rq)
            elif op[0] == "ratio":
                chunk = op[1]
                old = True
                if rq in pos:
                    t = ticket_of(pos[rq] // tile_rows, chunk)
                    if t < self.c_row + base:                # claimed before my snapshot: the record says what happened
                        if t not in rec:
                            return True                      # keep waiting for that tile
                        old = rec[t][1]
                for i in range(chunk * chunk_cols, (chunk + 1) * chunk_cols):
                    if old:
                        T[rq][i] = rowp[rq] / piv if i == lp else T[rq][i] + s[i] * rowp[rq]
                    out["col"][i] = T[rq][i]
            elif op[0] == "barrier_R":
                barrier["R"].add(self.h)
                if len(barrier["R"]) < H:
                    return True
            elif op[0] == "pub_col":
                self.c_col = word.or_col(NONE_COL if same_col else lpn)
            elif op[0] == "gather":
                r = op[1]
                if same_col:
                    out["rowp"][r] = rowp[r] / piv
                elif r == 0 or r == rq or r not in pos:
                    v = T[r][lpn]                            # final (stages 0 / R), or a row no tile touches
                    if r not in pos and r != 0 and r != rq:
                        v = v + s[lpn] * rowp[r]             # (a_pr == 0: the trivial update)
                    out["rowp"][r] = v
                else:
                    t = ticket_of(pos[r] // tile_rows, lpn // chunk_cols)
                    held = True
                    if t < self.c_col + base:
                        if t not in rec:
                            return True
                        held = rec[t][2]
                    out["rowp"][r] = T[r][lpn] + s[lpn] * rowp[r] if held else T[r][lpn]
            elif op[0] == "barrier_G":
                barrier["G"].add(self.h)
                if len(barrier["G"]) < H:
                    return True
            elif op[0] == "quiet":
                word.quiet = True
            self.todo.pop(0)
            return True

    barrier = {"R": set(), "G": set()}
    ctas = [Cta(g) for g in range(G)]
    helpers = [Helper(h) for h in range(H)]
    # helpers stream too once their chain is done: model them as extra CTAs that start late (no implicit tile)
    late = []
    actors = ctas + helpers
    guard = 0
    while actors:
        guard += 1
        assert guard < 200000, "protocol model deadlocked"
        a = rng.choice(actors)
        alive = a.step()
        if not alive:
            actors.remove(a)
            if isinstance(a, Helper):
                c = Cta(ntiles)                              # no implicit tile
                c.state = "claim"
                late.append(c)
                actors.append(c)

    # rows the list leaves out: the next update's helpers write the true pivot-column entry (stage 0 repair) -- here the
    # CURRENT pivot column of skipped rows must already be true because the previous flush / repair made it so
    for r in range(R):
        if r != 0 and r not in pos:
            T[r][lp] = rowp[r] / piv                         # = a_pr (zero) / pivot; what the repair writes

    assert out["b"] == new[0]
    assert out["col"] == new[rq]
    assert out["rowp"] == [new[r][lpn] for r in range(R)]
    for r in range(R):
        for i in range(m):
            if i != lpn or same_col:
                assert T[r][i] == new[r][i], (seed, r, i)
    for r in range(R):                                       # la_flush_kernel
        T[r][lpn] = out["rowp"][r]
    assert T == new


@pytest.mark.parametrize("block", range(8))
def test_lookahead_protocol_model_random_interleavings(block):
    for seed in range(block * 60, block * 60 + 60):
        run_trial(seed)
