"""Exhaustive interleaving check of the inter-GPU protocols (two-shot kernel, sharded step, staged exchange).

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


def explore_sharded(world, steps, grad_halves=2, gather_barrier=True, max_states=3_000_000):
    """Same exploration for the fused sharded step (k456_sharded_kernel): gradients staged in a double
    buffer, parameters SINGLE buffered in the symmetric arena.  Step k of rank r:
        stage grads(k) -> barrier -> reduce peers' grads(k), Adam: params[r] := k -> barrier -> read params[p] == k.
    The claim checked: no trailing barrier is needed, because rank r rewrites params[r] for step k+1 only after
    the first barrier of step k+1, which every peer reaches only after its gather of step k."""

    def program(k):
        e1, e2 = 2 * k + 1, 2 * k + 2
        pr = [("stage", k), ("arrive", e1)] + [("wait", (e1, p)) for p in range(world)]
        pr += [("reduce_read", (k, p)) for p in range(world)] + [("adam", k)]
        if gather_barrier:
            pr += [("arrive", e2)] + [("wait", (e2, p)) for p in range(world)]
        pr += [("gather_read", (k, p)) for p in range(world) ]
        return pr

    flat = [op for k in range(steps) for op in program(k)]
    n_ops = len(flat)
    init = (tuple([0] * world), tuple(tuple([-1] * grad_halves) for _ in range(world)), tuple([-1] * world),
            tuple(tuple(tuple([0] * world) for _ in range(2)) for _ in range(world)))
    seen, todo = {init}, deque([init])
    while todo:
        pcs, grads, params, flags = todo.popleft()
        moved = False
        for r in range(world):
            pc = pcs[r]
            if pc == n_ops:
                continue
            op, arg = flat[pc]
            ng, npar, nf = grads, params, flags
            if op == "stage":
                row = list(grads[r]); row[arg % grad_halves] = arg
                ng = grads[:r] + (tuple(row),) + grads[r + 1:]
            elif op == "adam":
                npar = params[:r] + (arg,) + params[r + 1:]
            elif op == "arrive":
                fl = [list(map(list, f)) for f in flags]
                for p in range(world):
                    fl[p][arg % 2][r] = arg
                nf = tuple(tuple(tuple(x) for x in f) for f in fl)
            elif op == "wait":
                e, p = arg
                if flags[r][e % 2][p] < e:
                    continue
            elif op == "reduce_read":
                k, p = arg
                if grads[p][k % grad_halves] != k:
                    return "rank %d reduces step %d but sees gradient version %d of rank %d" % (r, k, grads[p][k % grad_halves], p)
            elif op == "gather_read":
                k, p = arg
                if p != r and params[p] != k:
                    return "rank %d gathers step %d but sees parameter version %d of rank %d" % (r, k, params[p], p)
            moved = True
            nxt = (pcs[:r] + (pc + 1,) + pcs[r + 1:], ng, npar, nf)
            if nxt not in seen:
                seen.add(nxt)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
                todo.append(nxt)
        if not moved and any(pc != n_ops for pc in pcs):
            return "deadlock at program counters %r" % (pcs,)
    return None


@pytest.mark.parametrize("world,steps", [(2, 4), (3, 3)])
def test_fused_sharded_step_needs_no_trailing_barrier(world, steps):
    assert explore_sharded(world, steps) is None


def test_the_sharded_model_catches_a_missing_gather_barrier():
    bad = explore_sharded(2, 2, gather_barrier=False)
    assert bad is not None and "parameter version" in bad


def explore_staged(world, steps, chunks, halves=2, reuse_edge=True, max_states=4_000_000):
    """The staged exchange (csrc/b2d_staged.cuh): per rank three concurrent in-order streams coupled by two
    monotone flags per source rank and ONE local ordering edge —
        S  for every (step k, chunk c): [local: U has finished step k - halves]  stage chunk c into half k % halves;
           publish staged = idx
        X  wait staged[p] >= idx for ALL p (own included); read chunk c / own slice of every rank (must be the staged
           values of step k); PUSH the reduced slice into every rank's half; publish published = idx
        U  wait published[p] >= idx for ALL p; read the whole chunk from the OWN half (must be step k's reduced values)
    No kernel ever waits after it has published, and nothing else orders the streams of a rank (the event edges
    S->X and X->U of b2d.cu are implied by the flags, so the model is strictly more permissive than the code).
    Checked: every read sees the (kind, step) it expects; no reachable state is a deadlock."""
    idx = lambda k, c: k * chunks + c + 1
    prog = {
        "S": [x for k in range(steps) for c in range(chunks) for x in (("stage", (k, c)), ("pubA", idx(k, c)))],
        "X": [x for k in range(steps) for c in range(chunks) for x in (("waitA", idx(k, c)), ("exch", (k, c)), ("pubB", idx(k, c)))],
        "U": [x for k in range(steps) for c in range(chunks) for x in (("waitB", idx(k, c)), ("unstage", (k, c)))],
    }
    roles = ("S", "X", "U")
    empty = (-1, -1)
    mem0 = tuple(tuple(tuple(tuple([empty] * world) for _ in range(chunks)) for _ in range(halves)) for _ in range(world))
    zero = tuple(tuple([0] * world) for _ in range(world))
    init = (tuple((0, 0, 0) for _ in range(world)), mem0, zero, zero)
    seen, todo = {init}, deque([init])

    def freeze(m):
        return tuple(tuple(tuple(tuple(x) for x in ch) for ch in hf) for hf in m)

    def thaw(mem):
        return [[[list(x) for x in ch] for ch in hf] for hf in mem]

    while todo:
        pcs, mem, cntA, cntB = todo.popleft()
        moved = False
        done_all = True
        for r in range(world):
            for ri, role in enumerate(roles):
                pc = pcs[r][ri]
                if pc == len(prog[role]):
                    continue
                done_all = False
                op, arg = prog[role][pc]
                nmem, nA, nB = mem, cntA, cntB
                if op == "stage":
                    k, c = arg
                    if reuse_edge and k >= halves and pcs[r][2] < 2 * chunks * (k - halves + 1):
                        continue          # the half's previous write-back has not finished (reuse_ev in b2d.cu)
                    m = thaw(mem)
                    m[r][k % halves][c] = [(0, k)] * world
                    nmem = freeze(m)
                elif op == "pubA":
                    a = [list(x) for x in cntA]
                    for p in range(world):
                        a[p][r] = arg
                    nA = tuple(tuple(x) for x in a)
                elif op == "pubB":
                    b = [list(x) for x in cntB]
                    for p in range(world):
                        b[p][r] = arg
                    nB = tuple(tuple(x) for x in b)
                elif op == "waitA":
                    if any(cntA[r][p] < arg for p in range(world)):
                        continue
                elif op == "waitB":
                    if any(cntB[r][p] < arg for p in range(world)):
                        continue
                elif op == "exch":
                    k, c = arg
                    for p in range(world):
                        if mem[p][k % halves][c][r] != (0, k):
                            return "rank %d reduces (step %d, chunk %d) but rank %d holds %r" % (r, k, c, p, mem[p][k % halves][c][r])
                    m = thaw(mem)
                    for p in range(world):
                        m[p][k % halves][c][r] = (1, k)
                    nmem = freeze(m)
                elif op == "unstage":
                    k, c = arg
                    for sl in range(world):
                        if mem[r][k % halves][c][sl] != (1, k):
                            return "rank %d writes back (step %d, chunk %d) but slice %d holds %r" % (r, k, c, sl, mem[r][k % halves][c][sl])
                moved = True
                row = list(pcs[r]); row[ri] = pc + 1
                nxt = (pcs[:r] + (tuple(row),) + pcs[r + 1:], nmem, nA, nB)
                if nxt not in seen:
                    seen.add(nxt)
                    if len(seen) > max_states:
                        raise RuntimeError("state space larger than expected")
                    todo.append(nxt)
        if not moved and not done_all:
            return "deadlock at pcs %r" % (pcs,)
    return None


@pytest.mark.parametrize("world,steps,chunks", [(2, 3, 2), (2, 4, 1), (3, 3, 1)])
def test_staged_exchange_protocol_is_safe(world, steps, chunks):
    assert explore_staged(world, steps, chunks) is None


def test_staged_exchange_single_buffered_is_safe_too():
    """With the write-back edge even ONE half is safe (consecutive uses of a slot just serialise)."""
    assert explore_staged(2, 3, 2, halves=1) is None


def test_staged_model_catches_a_missing_reuse_edge():
    bad = explore_staged(2, 3, 1, reuse_edge=False)
    assert bad is not None and "holds" in bad


def explore_owner(world, steps, buckets, step_fence=True, max_states=4_000_000):
    """The sharded path (csrc/b2d_owner.cuh): per rank, per step k —
        S  for every reduce bucket b: stage b into the bucket's SINGLE staging region; publish staged = idx
        X  for every bucket: wait staged of ALL ranks; the owner side reads the bucket's staging region of every rank
           (must hold step k); after the last bucket: push the own parameter shard into every rank (version k);
           publish published = k + 1
        U  wait published of ALL ranks; now every rank's copy of every shard must be version k (the step returns)
    The staging regions have no double buffer: what fences their re-use is that S of step k+1 starts only after the rank's
    own U of step k (host stream order: backward follows optimizer.step()).  Checked: every read sees the step it
    expects, no deadlock; without that fence a violating schedule exists."""
    idx = lambda k, b: k * buckets + b + 1
    prog = {
        "S": [x for k in range(steps) for b in range(buckets) for x in (("stage", (k, b)), ("pubA", idx(k, b)))],
        "X": [x for k in range(steps)
              for x in ([y for b in range(buckets) for y in (("waitA", idx(k, b)), ("reduce", (k, b)))] + [("push", k), ("pubB", k + 1)])],
        "U": [x for k in range(steps) for x in (("waitB", k + 1), ("check", k))],
    }
    roles = ("S", "X", "U")
    stage0 = tuple(tuple([-1] * buckets) for _ in range(world))          # stage[rank][bucket] = step staged
    par0 = tuple(tuple([-1] * world) for _ in range(world))              # par[rank][owner] = version of owner's shard held by rank
    zero = tuple(tuple([0] * world) for _ in range(world))
    init = (tuple((0, 0, 0) for _ in range(world)), stage0, par0, zero, zero)
    seen, todo = {init}, deque([init])
    while todo:
        pcs, stage, par, cntA, cntB = todo.popleft()
        moved, done_all = False, True
        for r in range(world):
            for ri, role in enumerate(roles):
                pc = pcs[r][ri]
                if pc == len(prog[role]):
                    continue
                done_all = False
                op, arg = prog[role][pc]
                nstage, npar, nA, nB = stage, par, cntA, cntB
                if op == "stage":
                    k, b = arg
                    if step_fence and k >= 1 and pcs[r][2] < 2 * k:
                        continue          # backward of step k starts after optimizer.step() of step k-1 has returned
                    st = [list(x) for x in stage]
                    st[r][b] = k
                    nstage = tuple(tuple(x) for x in st)
                elif op == "pubA":
                    a = [list(x) for x in cntA]
                    for p in range(world):
                        a[p][r] = arg
                    nA = tuple(tuple(x) for x in a)
                elif op == "pubB":
                    bb = [list(x) for x in cntB]
                    for p in range(world):
                        bb[p][r] = arg
                    nB = tuple(tuple(x) for x in bb)
                elif op == "waitA":
                    if any(cntA[r][p] < arg for p in range(world)):
                        continue
                elif op == "waitB":
                    if any(cntB[r][p] < arg for p in range(world)):
                        continue
                elif op == "reduce":
                    k, b = arg
                    for p in range(world):
                        if stage[p][b] != k:
                            return "rank %d reduces bucket %d of step %d but rank %d staged step %r" % (r, b, k, p, stage[p][b])
                elif op == "push":
                    pr = [list(x) for x in par]
                    for p in range(world):
                        pr[p][r] = arg
                    npar = tuple(tuple(x) for x in pr)
                elif op == "check":
                    for o in range(world):
                        if par[r][o] != arg:
                            return "rank %d leaves step %d holding parameter version %r of owner %d" % (r, arg, par[r][o], o)
                moved = True
                row = list(pcs[r]); row[ri] = pc + 1
                nxt = (pcs[:r] + (tuple(row),) + pcs[r + 1:], nstage, npar, nA, nB)
                if nxt not in seen:
                    seen.add(nxt)
                    if len(seen) > max_states:
                        raise RuntimeError("state space larger than expected")
                    todo.append(nxt)
        if not moved and not done_all:
            return "deadlock at pcs %r" % (pcs,)
    return None


@pytest.mark.parametrize("world,steps,buckets", [(2, 3, 2), (3, 2, 2), (2, 4, 1)])
def test_owner_path_protocol_is_safe(world, steps, buckets):
    assert explore_owner(world, steps, buckets) is None


def test_owner_model_catches_a_missing_step_fence():
    """Single-buffered staging is only safe because backward k+1 follows optimizer.step() k: without it a fast rank
    re-stages a bucket a slow owner has not reduced yet."""
    bad = explore_owner(2, 2, 1, step_fence=False)
    assert bad is not None and ("staged step" in bad or "parameter version" in bad)
