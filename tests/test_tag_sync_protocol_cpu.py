"""Executable model of the persistent decode kernel's "tag sync" (csrc/decode_persistent.cu): CTAs as sequential programs, every value
that crosses CTAs as a (generation) cell of a double-buffered instance with a one-bit tag, a random scheduler that lets any CTA lag
arbitrarily far behind.  What the kernel's header argues in prose is checked here by exhaustion over random interleavings:

  * safety   - a reader that accepts a cell (its tag is the expected one) always holds exactly the generation it wanted: no stale value
               of generation c - 2, no value of generation c + 2 that a fast writer put there early (buffer re-use without a barrier);
  * liveness - the run finishes (no reader waits for a tag that can never appear).

The read / write sets per phase are the kernel's (linear_phase / attention_phase / mlp2_finalize / embed_row, ownership by part_range,
quarter_range and pair = blockIdx + k * grid), including CTAs that own nothing in a phase and skip its reads.  A negative control
(readers that do not look at the tag) must trip the checker, so the model has teeth."""
import random

import pytest


def part_range(bx, U, G):
    return (bx * U) // G, ((bx + 1) * U) // G


def quarter_range(bx, upq, G):
    cpq = G // 4
    q, j = bx // cpq, bx % cpq
    if q >= 4:
        return 0, 0
    return q * upq + (j * upq) // cpq, q * upq + ((j + 1) * upq) // cpq


def inst(c):
    return c & 1


def tag(c):
    return ((c >> 1) & 1) ^ 1


class Cell:
    """one tagged value (or group of values written together by one thread at one time)"""
    __slots__ = ("gen", "tag")

    def __init__(self):
        self.gen, self.tag = None, 0          # zeroed workspace: tag 0, no generation


class Model:
    def __init__(self, G, H, B, n_layers, n_steps, check_tags=True, seed=0):
        self.G, self.H, self.B, self.L, self.S = G, H, B, n_layers, n_steps
        self.check_tags = check_tags
        self.rng = random.Random(seed)
        d8 = 8 * H                              # row units of d = 64 H: d / 8
        self.Uq, self.U1, self.Uf = 3 * d8, 4 * d8, d8
        self.pairs = [(b, h) for b in range(B) for h in range(H)]
        # buffers: name -> [instance][cell]; cells are indexed by the producer's work item
        self.buf = {n: [[Cell() for _ in range(k)] for _ in range(2)] for n, k in
                    (("XF", self.Uf), ("QKV", self.Uq * B), ("X1F", len(self.pairs)), ("HF", self.U1), ("P2", 4 * self.Uf))}
        self.logits = [None] * self.G
        self.violations = []
        self.barrier_count, self.barrier_gen = 0, 0
        self.progs = [self.program(bx) for bx in range(G)]
        self.done = [False] * G

    # ---- primitive steps (generators yield after every memory operation so that the scheduler can interleave anywhere)
    def read(self, name, c, cells, who):
        want = tag(c)
        for i in cells:
            while True:
                cell = self.buf[name][inst(c)][i]
                yield
                if not self.check_tags or cell.tag == want:
                    if cell.gen != c:
                        self.violations.append((who, name, i, "wanted", c, "got", cell.gen))
                    break

    def write(self, name, c, cells):
        for i in cells:
            cell = self.buf[name][inst(c)][i]
            cell.gen, cell.tag = c, tag(c)
            yield

    def grid_barrier(self):
        gen = self.barrier_gen
        self.barrier_count += 1
        if self.barrier_count == self.G:
            self.barrier_count, self.barrier_gen = 0, gen + 1
        while self.barrier_gen == gen:
            yield

    # ---- one CTA
    def program(self, bx):
        G, H, B = self.G, self.H, self.B
        d8 = 8 * H
        q0, q1 = part_range(bx, self.Uq, G)
        m0, m1 = part_range(bx, self.U1, G)
        k0, k1 = quarter_range(bx, self.Uf, G)
        my_pairs = [i for i in range(len(self.pairs)) if i % G == bx]
        who = f"cta{bx}"
        if bx < B:
            yield from self.write("XF", -1, range(self.Uf) if B == 1 else range(bx, self.Uf, B))      # embedding of the first token (row bx: modelled as a share of the cells)
        yield from self.grid_barrier()
        c = 0
        for s in range(self.S):
            for _ in range(self.L):
                # QKV: statistics + the whole vector of the previous generation (also by pair owners without units: LN1 statistics)
                if q1 > q0 or my_pairs:
                    yield from self.read("XF", c - 1, range(self.Uf), who + " qkv")
                yield from self.write("QKV", c, [u * B + b for u in range(q0, q1) for b in range(B)])
                # attention: q / k / v units of the pair's head (8 units each), the previous residual segment; then X1
                for i in my_pairs:
                    b, h = self.pairs[i]
                    units = [kind * d8 + h * 8 + j for kind in range(3) for j in range(8)]
                    yield from self.read("QKV", c, [u * B + b for u in units], who + " att")
                    yield from self.read("XF", c - 1, range(h * 8, h * 8 + 8), who + " merge")
                yield from self.write("X1F", c, my_pairs)
                # MLP1: every pair's output (vector + LN2 partial sums)
                if m1 > m0:
                    yield from self.read("X1F", c, range(len(self.pairs)), who + " mlp1")
                yield from self.write("HF", c, range(m0, m1))
                # MLP2 partials of the CTA's K-quarter, then the finaliser of row unit bx
                if k1 > k0:
                    q = k0 // self.Uf
                    yield from self.read("HF", c, range(q * d8, (q + 1) * d8), who + " mlp2")
                yield from self.write("P2", c, range(k0, k1))
                if bx < self.Uf:
                    yield from self.read("P2", c, [q * self.Uf + bx for q in range(4)], who + " fin")
                    yield from self.read("X1F", c, [i for i, (b, h) in enumerate(self.pairs) if h == bx // 8], who + " fin x1")
                    yield from self.write("XF", c, [bx])
                c += 1
            # head (reads the last layer's vector), real barrier, sampling + embedding by CTAs b < B as generation c - 1, real barrier
            yield from self.read("XF", c - 1, range(self.Uf), who + " head")
            yield from self.grid_barrier()
            if bx < B:
                yield from self.write("XF", c - 1, range(self.Uf) if B == 1 else range(bx, self.Uf, B))
            yield from self.grid_barrier()
        self.done[bx] = True

    def run(self, max_events=3_000_000, lag=None):
        """random scheduler; `lag`: a CTA that is only scheduled with a small probability (arbitrarily slow reader / writer)"""
        live = list(range(self.G))
        n = 0
        while live:
            bx = self.rng.choice(live)
            if lag is not None and bx == lag and self.rng.random() < 0.9:
                continue
            try:
                next(self.progs[bx])
            except StopIteration:
                live.remove(bx)
            n += 1
            assert n < max_events, "no progress: a reader waits for a tag that never appears"
            if self.violations:
                break
        return self.violations


CASES = [  # G, H, B (G >= d / 8 = 8 H: every row unit has a finaliser CTA, as the launcher demands): ownership patterns incl. CTAs without
    # QKV / MLP units, without pairs, with several pairs, without a K-quarter (G % 4 != 0)
    (8, 1, 2), (16, 2, 3), (9, 1, 1), (16, 1, 2), (9, 1, 16), (17, 2, 1), (11, 1, 3),
]


@pytest.mark.parametrize("G,H,B", CASES)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tag_sync_is_safe_and_live_under_random_interleavings(G, H, B, seed):
    for lag in (None, seed % G, G - 1):
        m = Model(G, H, B, n_layers=5, n_steps=2, seed=seed * 7 + (0 if lag is None else lag + 1))
        assert m.run(lag=lag) == []
        assert all(m.done)


def test_the_model_catches_readers_that_ignore_the_tag():
    found = False
    for seed in range(6):
        m = Model(8, 1, 2, n_layers=5, n_steps=2, check_tags=False, seed=seed)
        if m.run(lag=seed % 8):
            found = True
            break
    assert found, "negative control: reads without the tag check must observe a wrong generation in some interleaving"
