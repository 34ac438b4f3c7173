"""Randomised-schedule model of the warp-level stage synchronisation of qme_tile_kernel (template flag V & 2,
lime_b200/csrc/qme_tile.cuh).  Not product code: it replays the PROTOCOL -- per-warp mbarriers indexed by the
stage parity, symmetric dependency sets, halo mbarriers with transaction counts armed by thread 0, halo pushes from
the first / last patch rows -- under random interleavings of the warps of a whole cluster and checks every shared-
memory read against the version (stage number) of the word it reads.  A read-after-write or write-after-read
violation, a wait that can never complete, or an mbarrier parity alias shows up as an assertion / deadlock here
instead of on the GPU.

    python tools/tile_sync_model.py [runs]
"""
import random
import sys


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, 'arrival on a barrier whose phase is already complete (count overflow)'
        self.tx += tx
        self.pending -= 1
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def test(self, parity):            # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def simulate(rng, P, chunk, C, TR, CB, shifts, colshift, nstages, obs):
    """shifts[p] = (p', d): the sandwich source of path p row t is path p' row t + d (|d| <= 1), column block
    colshift.  Cells are (path, t, cb) with t in 0..chunk+1 (0 and chunk+1 = halo rows)."""
    G = P * chunk // TR                       # row groups
    W = G * CB

    def patch(w):
        g, cb = divmod(w, CB)
        own0 = g * TR
        return own0 // chunk, 1 + own0 % chunk, cb     # path, first row t0, column block

    # --- what a warp reads per patch row r (list of cells), per the kernel's stage function
    def reads_of(w):
        p, t0, cb = patch(w)
        rows = []
        for r in range(TR):
            t = t0 + r
            cells = [(p, t - 1, cb), (p, t, cb), (p, t + 1, cb)]            # window
            cells += [(p, t, (cb - 1) % CB), (p, t, (cb + 1) % CB)]          # left / right neighbour columns
            if shifts is not None:
                pp, d = shifts[p]
                cells += [(pp, t + d, cb), (pp, t + d, (cb + colshift) % CB)]
            rows.append(cells)
        return rows

    def writer(cell):
        p, t, cb = cell
        if t == 0 or t == chunk + 1:
            return None
        return ((p * chunk + t - 1) // TR) * CB + cb

    reads = [reads_of(w) for w in range(W)]
    # host check of build_tile_host: a halo row may only be read by a patch that pushes in that direction
    for w in range(W):
        p, t0, cb = patch(w)
        for cells in reads[w]:
            for (pp, t, c2) in cells:
                if t == 0 and t0 != 1:
                    return 'rejected'
                if t == chunk + 1 and t0 + TR - 1 != chunk:
                    return 'rejected'
                if t < 0 or t > chunk + 1:
                    return 'rejected'

    # --- per CTA state
    class CTA:
        pass
    ctas = []
    for rank in range(C):
        c = CTA()
        c.rank = rank
        # version of every cell in the two buffers; buffer 0 holds the input of stage 0 (version 0)
        c.ver = [{(p, t, cb): (0 if b == 0 else -1) for p in range(P) for t in range(chunk + 2) for cb in range(CB)}
                 for b in range(2)]
        src = [set() for _ in range(W)]
        halo_read = [False] * W
        for w in range(W):
            for cells in reads[w]:
                for cell in cells:
                    o = writer(cell)
                    if o is None:
                        t = cell[1]
                        if (t == 0 and rank > 0) or (t == chunk + 1 and rank < C - 1):
                            halo_read[w] = True
                    else:
                        src[w].add(o)
        c.deps = []
        for w in range(W):
            readers = {v for v in range(W) if w in src[v]}
            c.deps.append((src[w] | readers) - {w})
        c.halo_wait = [(C > 1) and (halo_read[w] or w == 0) for w in range(W)]
        c.wbar = [[MBar(max(len(c.deps[w]), 1)) for _ in range(2)] for w in range(W)]
        c.hbar = [MBar(1), MBar(1)]
        c.halo_bytes = (P * CB if rank > 0 else 0) + (P * CB if rank < C - 1 else 0)
        c.state = [dict(stage=0, pc=0) for _ in range(W)]
        c.cta_bar = 0          # __syncthreads arrivals (observables)
        ctas.append(c)

    # program of one warp for one stage: list of micro-ops
    def program(c, w):
        p, t0, cb = patch(w)
        ops = []
        if w == 0 and C > 1:
            ops.append(('arm',))
        for r in range(TR):
            ops.append(('read', r))
            ops.append(('write', r))
            if r == 0 and t0 == 1 and c.rank > 0:
                ops.append(('push', r, c.rank - 1, chunk + 1))
            if r == TR - 1 and t0 + TR - 1 == chunk and c.rank < C - 1:
                ops.append(('push', r, c.rank + 1, 0))
        ops.append(('arrive',))
        if c.deps[w]:
            ops.append(('wait_own',))
        if c.halo_wait[w]:
            ops.append(('wait_halo',))
        ops.append(('endstage',))
        return ops

    progs = [[program(c, w) for w in range(W)] for c in ctas]
    total = C * W
    # some warps are much slower than others (a lagging warp is what breaks naive barrier reuse)
    speed = {(c.rank, w): rng.choice([1.0, 1.0, 0.2, 0.02, 5.0]) for c in ctas for w in range(W)}
    done = 0
    steps = 0
    while done < total:
        steps += 1
        assert steps < 5_000_000, 'livelock'
        runnable = []
        for c in ctas:
            for w in range(W):
                st = c.state[w]
                if st['stage'] >= nstages:
                    continue
                op = progs[c.rank][w][st['pc']]
                s = st['stage']
                if op[0] == 'wait_own' and not c.wbar[w][s & 1].test((s >> 1) & 1):
                    continue
                if op[0] == 'wait_halo' and not c.hbar[s & 1].test((s >> 1) & 1):
                    continue
                if op[0] == 'cta_sync' and False:
                    continue
                runnable.append((c, w))
        assert runnable, 'deadlock at ' + str([[st for st in c.state] for c in ctas])
        c, w = rng.choices(runnable, weights=[speed[(cc.rank, ww)] for cc, ww in runnable])[0]
        st = c.state[w]
        s = st['stage']
        op = progs[c.rank][w][st['pc']]
        p, t0, cb = patch(w)
        bi, bo = s & 1, (s + 1) & 1
        if op[0] == 'arm':
            c.hbar[s & 1].arrive(tx=c.halo_bytes)
        elif op[0] == 'read':
            for cell in reads[w][op[1]]:
                t = cell[1]
                o = writer(cell)
                boundary = o is None and not ((t == 0 and c.rank > 0) or (t == chunk + 1 and c.rank < C - 1))
                if boundary:
                    continue                       # zero row at the end of the chain: never written
                v = c.ver[bi][cell]
                assert v == s, 'hazard: CTA %d warp %d stage %d reads %s with version %d' % (c.rank, w, s, cell, v)
        elif op[0] == 'write':
            cell = (p, t0 + op[1], cb)
            assert c.ver[bo][cell] in (s - 1, -1), 'write order'
            c.ver[bo][cell] = s + 1
        elif op[0] == 'push':
            dst = ctas[op[2]]
            cell = (p, op[3], cb)
            dst.ver[bo][cell] = s + 1
            dst.hbar[s & 1].complete_tx(1)
        elif op[0] == 'arrive':
            for d in c.deps[w]:
                c.wbar[d][s & 1].arrive()
        elif op[0] == 'endstage':
            st['stage'] += 1
            st['pc'] = -1
            if st['stage'] >= nstages:
                done += 1
        st['pc'] += 1
    return 'ok'


def main(runs):
    rng = random.Random(1234)
    ok = rej = 0
    for it in range(runs):
        P = rng.choice([1, 2, 2, 4])
        TR = 4
        CB = rng.choice([1, 2])
        C = rng.choice([1, 2, 4, 4])
        chunk = TR * rng.choice([1, 2, 4])
        if P * chunk // TR * CB > 16:
            continue
        if rng.random() < 0.2:
            shifts = None
        else:
            shifts = {p: (rng.randrange(P), rng.choice([-1, 0, 1])) for p in range(P)}
        r = simulate(rng, P, chunk, C, TR, CB, shifts, rng.randrange(CB), nstages=rng.choice([4, 8, 12]), obs=False)
        if r == 'ok':
            ok += 1
        else:
            rej += 1
    print('tile_sync_model: %d random geometries x schedules ok, %d rejected by the host-side halo rule' % (ok, rej))
    return ok


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
