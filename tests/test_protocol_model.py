"""Exhaustive interleaving check of the inter-GPU protocol of the two-shot kernel.

The kernels synchronise same-index blocks of the W ranks with epoch flags and avoid a trailing barrier by
alternating two staging halves per bucket slot.  That argument is easy to get subtly wrong, and a GPU test
only samples a few schedules; this model explores EVERY interleaving of the ranks' atomic actions for a few
consecutive steps (sequentially consistent memory — what the system-scope fences around the flags give)
and checks that each read observes exactly the version it is meant to.

Modelled per rank and step k (one block; blocks never interact):
    stage    write own stage[half(k)] := k
    arrive   peers' flag[e&1][me] := e                        (epoch e = 2k+1)
    wait     until own flag[e&1][p] == e for every peer p
    reduce   read stage[half(k)] of every peer   (must be k);  own result[half(k)] := k
    arrive/wait at epoch e+1
    gather   read result[half(k)] of every peer  (must be k)
Also checked: no reachable state is a deadlock.  The same machinery shows that the checker is not vacuous:
with ONE staging half (no double buffering, no trailing barrier) it finds a violating schedule.
"""
from collections import deque

import pytest


def explore(world, steps, halves=2, flag_sets=2, max_states=3_000_000):
    """BFS over all interleavings. Returns None when every schedule is safe, else a description.

    Memory of rank p: mem[p][half][slice] = (kind, step) with kind 0 = staged gradient, 1 = reduced result —
    the reduced slice lives IN the staging buffer (own slice region), exactly as in the kernel, so a
    re-stage of the same half destroys it."""

    def program(k):
        e1, e2 = 2 * k + 1, 2 * k + 2
        pr = [("stage", k), ("arrive", e1)]
        pr += [("wait", (e1, p)) for p in range(world)]
        pr += [("reduce_read", (k, p)) for p in range(world)]
        pr += [("result", k), ("arrive", e2)]
        pr += [("wait", (e2, p)) for p in range(world)]
        pr += [("gather_read", (k, p)) for p in range(world)]
        return pr

    flat = [op for k in range(steps) for op in program(k)]
    n_ops = len(flat)
    empty = (-1, -1)
    init = (tuple([0] * world),
            tuple(tuple(tuple([empty] * world) for _ in range(halves)) for _ in range(world)),
            tuple(tuple(tuple([0] * world) for _ in range(flag_sets)) for _ in range(world)))
    seen = {init}
    todo = deque([init])
    while todo:
        pcs, mem, flags = todo.popleft()
        moved = False
        for r in range(world):
            pc = pcs[r]
            if pc == n_ops:
                continue
            op, arg = flat[pc]
            nmem, nflags = mem, flags
            if op == "stage":
                m = [list(map(list, h)) for h in mem]
                m[r][arg % halves] = [(0, arg)] * world
                nmem = tuple(tuple(tuple(sl) for sl in h) for h in m)
            elif op == "result":
                m = [list(map(list, h)) for h in mem]
                m[r][arg % halves][r] = (1, arg)
                nmem = tuple(tuple(tuple(sl) for sl in h) for h in m)
            elif op == "arrive":
                fl = [list(map(list, f)) for f in flags]
                for p in range(world):
                    fl[p][arg % flag_sets][r] = arg
                nflags = tuple(tuple(tuple(x) for x in f) for f in fl)
            elif op == "wait":
                e, p = arg
                if flags[r][e % flag_sets][p] < e:
                    continue                             # blocked: the peer has not arrived at epoch e yet
            elif op == "reduce_read":
                k, p = arg
                if mem[p][k % halves][r] != (0, k):
                    return "rank %d reduces step %d but rank %d's slice holds %r" % (r, k, p, mem[p][k % halves][r])
            elif op == "gather_read":
                k, p = arg
                if mem[p][k % halves][p] != (1, k):
                    return "rank %d gathers step %d but rank %d's slice holds %r" % (r, k, p, mem[p][k % halves][p])
            moved = True
            nxt = (pcs[:r] + (pc + 1,) + pcs[r + 1:], nmem, nflags)
            if nxt not in seen:
                seen.add(nxt)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
                todo.append(nxt)
        if not moved and any(pc != n_ops for pc in pcs):
            return "deadlock at program counters %r" % (pcs,)
    return None


@pytest.mark.parametrize("world,steps", [(2, 4), (3, 3)])
def test_double_buffered_protocol_is_safe_under_every_interleaving(world, steps):
    assert explore(world, steps, halves=2, flag_sets=2) is None


def test_the_model_catches_a_missing_double_buffer():
    """One staging half and no trailing barrier: a fast rank re-stages step k+1 while a slow peer still gathers step k."""
    bad = explore(2, 3, halves=1, flag_sets=2)
    assert bad is not None and "gathers step" in bad and "holds (0," in bad   # a re-staged gradient where a result should be


def test_monotone_epochs_make_one_flag_set_enough_for_waiting():
    """With `>=` comparison a single flag set never blocks forever nor lets a rank through early: the data
    checks still hold (this is the scheme the split arrive/wait helpers rely on)."""
    assert explore(2, 3, halves=2, flag_sets=1) is None
