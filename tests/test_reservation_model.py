"""An executable model of how insert_kernel (kmcex_b200/csrc/kmx_build.cu) reproduces the reference's SEQUENTIAL
greedy insert (kmodel.hpp:543-555, 590-622) in parallel -- claims, reservations, classic two-barrier iterations and
merged reserve/commit passes -- run against the sequential loop on small random instances built to collide.

The model is not the CUDA code; it is the algorithm DESIGN.md section 3.3 describes, with the freedom the GPU has made
explicit: inside a phase the items act in an arbitrary order, a cell read may see any subset of the commits of the
same phase that happened before it, claim bitmaps and reservation tables alias positions.  Whatever the interleaving,
the set of accepted items and the final (tag, value) state must equal the sequential result."""
import random

H = 4                      # hashes per item (the real path uses n_hash = 7)
EPOCH_MAX = (1 << 14) - 1


def make_items(rng, n, n_pos):
    """items: (positions[H], wanted bits[H]); few positions so that items collide constantly"""
    return [([rng.randrange(n_pos) for _ in range(H)], [rng.randrange(2) for _ in range(H)]) for _ in range(n)]


def sequential(items, n_pos, tag=None, val=None):
    """insert_to_array, one item after the other (kmodel.hpp:590-622), on top of an existing state"""
    tag, val = list(tag or [0] * n_pos), list(val or [0] * n_pos)
    accepted = []
    for pos, want in items:
        ok = all(not tag[p] or val[p] == w for p, w in zip(pos, want))
        accepted.append(ok)
        if ok:
            for p, w in zip(pos, want):          # an item whose own hashes collide with different bits: checked against the
                tag[p] = 1                       # pre-item state, then written in hash order (kmodel.hpp:603-618)
                val[p] |= w
    return accepted, tag, val


class Parallel:
    def __init__(self, items, n_pos, rng, claim_bits, resv_slots, merged, tag=None, val=None):
        self.items, self.n_pos, self.rng, self.merged = items, n_pos, rng, merged
        self.tag, self.val = list(tag or [0] * n_pos), list(val or [0] * n_pos)
        self.claim_bits, self.resv_slots = claim_bits, resv_slots
        self.state = [None] * len(items)          # None undecided, True accepted, False rejected
        self.untagged = [0] * len(items)
        self.need = [0] * len(items)

    def read(self, i):
        pos, want = self.items[i]
        conflict, untagged = False, 0
        for j, (p, w) in enumerate(zip(pos, want)):
            if self.tag[p] and self.val[p] != w:
                conflict = True
            if not self.tag[p]:
                untagged |= 1 << j
        return conflict, untagged

    def commit(self, i, untagged):
        pos, want = self.items[i]
        for j, (p, w) in enumerate(zip(pos, want)):
            if (untagged >> j) & 1:               # one atomic OR of tag + value per untagged position
                self.tag[p] = 1
                self.val[p] |= w
        self.state[i] = True

    def shuffled(self, ids):
        ids = list(ids)
        self.rng.shuffle(ids)
        return ids

    def run(self):
        items, n = self.items, len(self.items)
        claim = [[0] * self.claim_bits for _ in range(2)]
        # phase 0: reject on the committed state (what earlier buckets / batches left in the array) or claim (position, want)
        for i in self.shuffled(range(n)):
            conflict, untagged = self.read(i)
            if conflict:
                self.state[i] = False
                continue
            self.untagged[i] = untagged
            for j, (p, w) in enumerate(zip(*items[i])):
                if (untagged >> j) & 1:
                    claim[w][p % self.claim_bits] = 1
        # phase 1: uncontested items commit, the others reserve; commits interleave with the other items' work
        tables = [[[(EPOCH_MAX << 18) | 0x3FFFF] * (2 * self.resv_slots)], [[(EPOCH_MAX << 18) | 0x3FFFF] * (2 * self.resv_slots)]]
        tables = [t[0] for t in tables]
        epoch = 0
        key_hi = (EPOCH_MAX - epoch) << 18
        contested = []
        for i in self.shuffled(j for j in range(n) if self.state[j] is None):
            pos, want = items[i]
            need = 0
            for j, (p, w) in enumerate(zip(pos, want)):
                if (self.untagged[i] >> j) & 1 and claim[w ^ 1][p % self.claim_bits]:
                    need |= 1 << j
            if need == 0:
                self.commit(i, self.untagged[i])
            else:
                self.need[i] = need
                self.reserve(tables[0], i, need, key_hi)
                contested.append(i)
        if self.merged:
            self.merged_passes(tables, contested, epoch)
        else:
            self.classic_iterations(tables[0], contested, epoch)
        return [bool(s) for s in self.state], self.tag, self.val

    def reserve(self, table, i, need, key_hi):
        for j, (p, w) in enumerate(zip(*self.items[i])):
            if (need >> j) & 1:
                slot = 2 * (p % self.resv_slots) + w
                table[slot] = min(table[slot], key_hi | i)

    def holds(self, table, i, need, key_hi):
        return all(table[2 * (p % self.resv_slots) + (w ^ 1)] >= (key_hi | i)
                   for j, (p, w) in enumerate(zip(*self.items[i])) if (need >> j) & 1)

    def classic_iterations(self, table, todo, epoch):
        key_hi = (EPOCH_MAX - epoch) << 18
        first = True
        while todo:
            if not first:                          # first half: re-read, reject / shrink the contested set, reserve
                epoch += 1
                key_hi = (EPOCH_MAX - epoch) << 18
                for i in self.shuffled(todo):
                    conflict, untagged = self.read(i)
                    if conflict:
                        self.state[i] = False
                        continue
                    self.untagged[i] = untagged
                    self.need[i] &= untagged
                    if self.need[i] == 0:
                        self.commit(i, untagged)
                    else:
                        self.reserve(table, i, self.need[i], key_hi)
            first = False
            nxt = []                               # second half, after a barrier: holders commit
            for i in self.shuffled(j for j in todo if self.state[j] is None):
                if self.holds(table, i, self.need[i], key_hi):
                    self.commit(i, self.untagged[i])
                else:
                    nxt.append(i)
            assert len(nxt) < len(todo) or not todo
            todo = nxt

    def merged_passes(self, tables, todo, epoch):
        tab, fresh = 0, True
        key_hi = (EPOCH_MAX - epoch) << 18
        passes = 0
        while todo:
            key_prev = key_hi
            epoch += 1
            key_hi = (EPOCH_MAX - epoch) << 18
            # every item reads its cells at some moment of the pass and acts at a later one; the events of all items
            # interleave arbitrarily (reads may or may not see the commits of the same pass)
            events = [(i, 0) for i in todo] + [(i, 1) for i in todo]
            self.rng.shuffle(events)
            seen_read, snapshot, nxt = set(), {}, []
            pending_act = set()
            for i, kind in events:
                if i not in seen_read:             # the first event of an item is its read, the second its action
                    seen_read.add(i)
                    snapshot[i] = (False, self.untagged[i]) if fresh else self.read(i)
                    pending_act.add(i)
                    continue
                pending_act.discard(i)
                conflict, untagged = snapshot[i]
                if conflict:
                    self.state[i] = False
                    continue
                need = self.need[i] & untagged
                if need == 0 or self.holds(tables[tab], i, need, key_prev):
                    self.commit(i, untagged)
                else:
                    self.reserve(tables[tab ^ 1], i, need, key_hi)
                    self.untagged[i], self.need[i] = untagged, need
                    nxt.append(i)
            passes += 1
            assert passes < 10 * len(self.items) + 10, "no progress"
            todo, tab, fresh = nxt, tab ^ 1, False


def check(seed, merged):
    rng = random.Random(seed)
    n = rng.randrange(1, 40)
    n_pos = rng.choice([3, 6, 12, 40])
    items = make_items(rng, n, n_pos)
    # what earlier buckets and batches left in the array (every other seed starts from an empty one)
    _, tag0, val0 = sequential(make_items(rng, rng.randrange(0, 6) if seed & 1 else 0, n_pos), n_pos)
    want = sequential(items, n_pos, tag0, val0)
    got = Parallel(items, n_pos, rng, claim_bits=rng.choice([1, 2, 5, 64]), resv_slots=rng.choice([1, 2, 7, 64]), merged=merged,
                   tag=tag0, val=val0).run()
    assert got == want, (seed, merged, items)


def test_classic_iterations_equal_the_sequential_insert():
    for seed in range(3000):
        check(seed, merged=False)


def test_merged_passes_equal_the_sequential_insert():
    for seed in range(3000):
        check(seed, merged=True)
