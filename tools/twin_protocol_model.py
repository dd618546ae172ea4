"""Executable model of the synchronisation protocol of fresco_attn_twin_kernel (fresco_b200/csrc/attn_tcgen05.cu).

The kernel cannot run without a GPU; its mbarrier protocol can be checked without one.  This file restates, agent by
agent, WHICH barrier every role waits on / arrives at and with WHICH phase parity -- the same expressions as the CUDA
source (cited inline) -- and runs the agents under a random interleaving with asynchronous engines (TMA completions, the
tensor core executing MMAs in issue order, tcgen05.commit arrivals).  It reports
  * a deadlock (nobody can make progress: a lost arrival, or a parity wait that aliased onto a later phase),
  * a data hazard: score region overwritten before its P was consumed, P V reading a region that does not hold that
    tile's P, a K/V ring stage refilled before its four readers are done, P V issued before the ones column of its V
    tile was patched, O rescaled while a P V of that query tile may still be in flight.
tests/test_twin_protocol_model.py sweeps tile counts, region counts (2 | 3), ring depths and threads per row.

mbarrier semantics modelled: `count` arrivals complete the current phase and flip the phase bit; `wait(parity)` passes
iff the phase with that parity has completed, i.e. the barrier's current phase bit differs from `parity`.  A barrier that
completes twice before a waiter looks is therefore seen as "not yet" -- exactly the aliasing hazard of parity waits.
"""
from __future__ import annotations

import random


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def done(self, parity):
        return self.phase != parity


class Hazard(AssertionError):
    pass


def simulate(n_tiles, nb=3, stages=4, split=1, fold=True, rescale_prob=0.2, seed=0, max_steps=2_000_000, final_on_pv=False):
    """One random interleaving.  Returns the number of scheduler steps; raises Hazard / RuntimeError('deadlock')."""
    rng = random.Random(seed)
    ST, NB = stages, nb
    n_total = 2 * n_tiles
    bar_q = Bar(1)
    kv_full = [Bar(1) for _ in range(ST)]
    kv_empty = [Bar(1) for _ in range(ST)]
    bar_s = [Bar(1) for _ in range(NB)]
    bar_p = [Bar(4 * split) for _ in range(NB)]             # mbar_init(bar_p + k, 4 * SPLIT)
    bar_pv = [Bar(1), Bar(1)]
    bar_vp = [Bar(2) for _ in range(ST)]                    # two patch warps
    bar_fin = [Bar(1), Bar(1)]                              # the last P V of a query tile; completes once
    final_wait_on_pv = final_on_pv                          # the pre-fix protocol (epilogue on bar_pv), kept to show the model catches it

    # ---- shared state for the hazard checks
    region = [None] * NB                                    # ("S", n) | ("P", n, parts written) | None
    stage_tile = [None] * ST                                # key tile held by a ring stage
    stage_reads = [0] * ST                                  # QK(2t), QK(2t+1), PV(2t), PV(2t+1) executed
    stage_patched = [False] * ST
    pv_executed = [0, 0]                                    # P V MMAs of query tile x executed so far
    pv_issued = [0, 0]
    loaded = {}                                             # (n, part) -> softmax part has its scores in registers

    tensor_q = []                                           # in-order queue of the tensor core: ("qk", n) | ("pv", n) | ("commit", bar)
    tma_q = []                                              # outstanding TMA loads, complete in any order: ("q",) | ("kv", t)

    def wait(bar, parity):
        while not bar.done(parity):
            yield "blocked"

    # ------------------------------------------------------------------ TMA producer (attn_tcgen05.cu: "TMA producer")
    def tma_producer():
        tma_q.append(("q",))
        yield "step"
        for t in range(n_tiles):
            st = t % ST
            if t >= ST:
                yield from wait(kv_empty[st], ((t // ST) - 1) & 1)
            if stage_tile[st] is not None and stage_reads[st] != 4:
                raise Hazard("ring stage %d refilled with tile %d while tile %s has %d of 4 reads" % (st, t, stage_tile[st], stage_reads[st]))
            tma_q.append(("kv", t))
            yield "step"

    # ------------------------------------------------------------------ TMA engine: loads land asynchronously
    def tma_engine():
        done = 0
        while done < n_tiles + 1:
            if not tma_q:
                yield "blocked"
                continue
            op = tma_q.pop(rng.randrange(len(tma_q)))
            if op[0] == "q":
                bar_q.arrive()
            else:
                t = op[1]
                st = t % ST
                stage_tile[st], stage_reads[st], stage_patched[st] = t, 0, False
                kv_full[st].arrive()
            done += 1
            yield "step"

    # ------------------------------------------------------------------ V patch warps (FOLD)
    def patcher():
        for t in range(n_tiles):
            st = t % ST
            yield from wait(kv_full[st], (t // ST) & 1)
            assert stage_tile[st] == t, ("patch of the wrong tile", st, t, stage_tile[st])
            bar_vp[st].arrive()                             # first patch warp
            yield "step"
            stage_patched[st] = True
            bar_vp[st].arrive()                             # second patch warp: the phase completes
            yield "step"

    # ------------------------------------------------------------------ MMA issuer ("the one MMA issuer")
    def issuer():
        def issue_qk(n):
            x, t = n & 1, n >> 1
            if x == 0:
                yield from wait(kv_full[t % ST], (t // ST) & 1)
            tensor_q.append(("qk", n))
            tensor_q.append(("commit", bar_s[n % NB]))
            yield "step"

        yield from wait(bar_q, 0)
        for n in range(min(NB, n_total)):
            yield from issue_qk(n)
        for n in range(n_total):
            yield from wait(bar_p[n % NB], (n // NB) & 1)
            if fold and not (n & 1):
                yield from wait(bar_vp[(n >> 1) % ST], ((n >> 1) // ST) & 1)
            tensor_q.append(("pv", n))
            pv_issued[n & 1] += 1
            tensor_q.append(("commit", bar_pv[n & 1]))
            if n + 2 >= n_total:
                tensor_q.append(("commit", bar_fin[n & 1]))
            if n & 1:
                tensor_q.append(("commit", kv_empty[(n >> 1) % ST]))
            yield "step"
            if n + NB < n_total:
                yield from issue_qk(n + NB)

    # ------------------------------------------------------------------ tensor core: executes in issue order
    def tensor_core():
        executed = 0
        total = None
        while True:
            if not tensor_q:
                if all(f for f in finished_flags["issuer"]):
                    return
                yield "blocked"
                continue
            op = tensor_q.pop(0)
            if op[0] == "qk":
                n = op[1]
                t, k = n >> 1, n % NB
                if region[k] is not None and not (region[k][0] == "consumed"):
                    raise Hazard("Q K^T of score tile %d overwrites region %d holding %r" % (n, k, region[k]))
                if stage_tile[t % ST] != t:
                    raise Hazard("Q K^T of score tile %d reads stage %d holding tile %r" % (n, t % ST, stage_tile[t % ST]))
                stage_reads[t % ST] += 1
                region[k] = ("S", n)
            elif op[0] == "pv":
                n = op[1]
                t, k, x = n >> 1, n % NB, n & 1
                if region[k] != ("P", n, split):
                    raise Hazard("P V of score tile %d reads region %d holding %r" % (n, k, region[k]))
                if stage_tile[t % ST] != t:
                    raise Hazard("P V of score tile %d reads stage %d holding tile %r" % (n, t % ST, stage_tile[t % ST]))
                if fold and not stage_patched[t % ST]:
                    raise Hazard("P V of score tile %d before the ones column of V(%d)" % (n, t))
                stage_reads[t % ST] += 1
                pv_executed[x] += 1
                region[k] = ("consumed",)
            else:
                op[1].arrive()
            executed += 1
            yield "step"

    # ------------------------------------------------------------------ softmax warp(s) of query tile x, key part `part`
    def softmax(x, part):
        m_set = False
        buf, ph = x % NB, 0
        for j in range(n_tiles):
            n = 2 * j + x
            assert buf == n % NB and ph == (n // NB) & 1, "buf / ph update rule"
            yield from wait(bar_s[buf], ph)
            if region[buf] is None or region[buf][1] != n or region[buf][0] not in ("S", "P"):
                raise Hazard("softmax of score tile %d finds region %d holding %r" % (n, buf, region[buf]))
            loaded[(n, part)] = True                        # tcgen05.ld of this part's columns
            yield "step"
            if j > 0 and rng.random() < rescale_prob:       # rare path: rescale O_x (needs P V_x(j-1) retired)
                yield from wait(bar_pv[x], (j - 1) & 1)
                if pv_executed[x] < j:
                    raise Hazard("O_%d rescaled at tile %d with %d P V executed" % (x, j, pv_executed[x]))
                if pv_issued[x] > j:
                    raise Hazard("O_%d rescaled at tile %d while P V %d may be in flight" % (x, j, pv_issued[x] - 1))
                yield "step"
            # P over the start of this part's OWN score columns (split 1: [0,64) of the region; split 2: [0,32) / [64,96))
            cur = region[buf]
            if cur[0] == "S":
                region[buf] = ("P", n, 1)
            else:
                region[buf] = ("P", n, cur[2] + 1)
            yield "step"
            for _ in range(4):                              # one elected arrival per warp (four row quarters)
                bar_p[buf].arrive()
            if NB == 3:
                if buf == 0:
                    buf = 2
                else:
                    buf -= 1
                    ph ^= 1
            else:
                ph ^= 1
            yield "step"
        if final_wait_on_pv:
            yield from wait(bar_pv[x], (n_tiles - 1) & 1)   # (the protocol before the fix)
        else:
            yield from wait(bar_fin[x], 0)                  # epilogue
        if pv_executed[x] != n_tiles:
            raise Hazard("epilogue of query tile %d after %d of %d P V" % (x, pv_executed[x], n_tiles))

    finished_flags = {"issuer": [False]}
    agents = {"tma": tma_producer(), "tma_engine": tma_engine(), "issuer": issuer(), "tensor": tensor_core()}
    if fold:
        agents["patch"] = patcher()
    for x in range(2):
        for part in range(split):
            agents["softmax%d.%d" % (x, part)] = softmax(x, part)
    alive = dict(agents)
    steps = 0
    blocked_streak = 0
    while alive:
        name = rng.choice(list(alive))
        try:
            r = next(alive[name])
        except StopIteration:
            del alive[name]
            if name == "issuer":
                finished_flags["issuer"][0] = True
            blocked_streak = 0
            continue
        steps += 1
        if r == "blocked":
            blocked_streak += 1
            if blocked_streak > 50 * len(alive) + 200:
                # everybody looked and nobody moved: confirm by polling each agent once more in order
                progress = False
                for nm in list(alive):
                    try:
                        if next(alive[nm]) != "blocked":
                            progress = True
                            break
                    except StopIteration:
                        del alive[nm]
                        if nm == "issuer":
                            finished_flags["issuer"][0] = True
                        progress = True
                        break
                if not progress:
                    raise RuntimeError("deadlock with %s alive after %d steps" % (sorted(alive), steps))
                blocked_streak = 0
        else:
            blocked_streak = 0
        if steps > max_steps:
            raise RuntimeError("no termination after %d steps" % steps)
    return steps


if __name__ == "__main__":
    for nb, st, split, fold in [(3, 4, 1, True), (3, 4, 2, True), (2, 2, 1, False), (2, 4, 2, False)]:
        for n_tiles in (1, 2, 3, 4, 7, 16):
            for seed in range(5):
                simulate(n_tiles, nb=nb, stages=st, split=split, fold=fold, seed=seed)
        print("regions %d, ring %d, threads/row %d, fold %s: ok" % (nb, st, split, fold))
