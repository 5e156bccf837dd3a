"""Explicit-state model check (CPU) of the mbarrier protocols of the tcgen05 MLP forward chain
(esr_nerf_b200/csrc/mlp_tc.cu:k_mlp_fwd_tc; the data-gradient chain k_mlp_dgrad_tc has the same structure with the
d_x accumulator in place of the output layer).

Three protocols are explored over EVERY interleaving of the MMA issuer thread, the epilogue warps and the (in-order,
asynchronous) tensor-core engine, for a CTA that processes several tiles:

  * "per_tile"  — the default build: every tile starts behind a __syncthreads, one mbarrier `bar` takes every commit;
  * "ovl_own"   — the opt-in tile overlap (template parameter OVL, ESR_MLP_TILE_OVERLAP=1): the next tile's layer-0 MMA
                  is issued behind this tile's output-layer MMA once all warps have arrived on `bar_x`, and commits to
                  its OWN mbarrier `bar_first`;
  * "ovl_shared" — the variant withdrawn at the end of round 1: the same overlap with the layer-0 commit on `bar`.

Model: an mbarrier is (completed phases, pending arrivals); `try_wait.parity P` succeeds iff the parity of the phase in
progress differs from P — so a waiter that sleeps through TWO completions sees its own parity again and never wakes.
Every TMEM / shared-memory region carries the (tile, layer) tag of its last writer; a reader asserts the tag it expects,
which catches an accumulator overwritten before a slow warp has read it and an operand consumed before it is complete.
The checker must prove the first two protocols free of deadlock and tag violations, and must FIND the hang of the third
(the diagnosis behind DESIGN.md §4 "Tile overlap (withdrawn)")."""
from collections import deque

import pytest

W = 2            # epilogue warps in the model (16 in the kernel: the protocol is symmetric in them)


def build_programs(proto: str, n_tiles: int, nh: int):
    """op lists of the issuer (thread 0) and the W epilogue warps (threads 1..W), with the parities each thread's
    toggling phase variables take at every wait (static: the control flow does not depend on data)"""
    ovl = proto != "per_tile"
    first_bar = "bar_first" if proto == "ovl_own" else "bar"
    d = lambda i: f"D{i & 1}"

    issuer, cphase, xphase, n_sync = [], 0, 0, 0
    mma_id = 0

    def layer0(tile):
        nonlocal mma_id
        issuer.append(("issue", mma_id, ("x", tile), d(0), (tile, 0), first_bar))
        mma_id += 1

    for t in range(n_tiles):
        more = t + 1 < n_tiles
        if not ovl or t == 0:
            issuer.append(("sync", n_sync))
            n_sync += 1
            layer0(t)
        for l in range(nh):                       # layer l + 1 (l + 1 == nh: the output layer) behind the chunk barrier
            issuer.append(("wait", "chunk", cphase))
            issuer.append(("issue", mma_id, ("A", (t, l)), d(l + 1), (t, l + 1), "bar"))
            mma_id += 1
            cphase ^= 1
        if ovl and more:
            issuer.append(("wait", "bar_x", xphase))
            xphase ^= 1
            layer0(t + 1)

    warps = []
    for w in range(W):
        ops, phase, fphase, n_sync = [("loadx", w, 0)], 0, 0, 0
        for t in range(n_tiles):
            more = t + 1 < n_tiles
            if not ovl or t == 0:
                ops.append(("sync", n_sync))
                n_sync += 1
            for l in range(nh):
                if l == 0 and proto == "ovl_own":
                    ops.append(("wait", "bar_first", fphase))
                    fphase ^= 1
                else:
                    ops.append(("wait", "bar", phase))
                    phase ^= 1
                if l == 0 and more:
                    ops.append(("loadx", w, t + 1))          # the x tile is free: prefetch the next one
                ops.append(("read", d(l), (t, l)))
                ops.append(("writeA", w, (t, l)))
                ops.append(("arrive", "chunk"))
            if ovl and more:
                ops.append(("arrive", "bar_x"))
            ops.append(("wait", "bar", phase))
            phase ^= 1
            ops.append(("read", d(nh), (t, nh)))
        warps.append(ops)
    return [issuer] + warps


BARS = {"bar": 1, "bar_first": 1, "bar_x": W, "chunk": W}
BAR_NAMES = sorted(BARS)


def explore(proto: str, n_tiles: int, nh: int, limit: int = 4_000_000):
    """BFS over all interleavings -> (states, deadlocks, violations)"""
    progs = build_programs(proto, n_tiles, nh)
    mmas = [op for op in progs[0] if op[0] == "issue"]
    issued_before = []                                       # number of MMAs issued once the issuer's pc has passed op i
    n = 0
    for op in progs[0]:
        issued_before.append(n)
        n += op[0] == "issue"
    issued_before.append(n)
    n_threads = len(progs)
    n_syncs = max((op[1] for p in progs for op in p if op[0] == "sync"), default=-1) + 1

    # state: pcs, bars ((done, pending) per name), tags (D0, D1, x[w]..., A[w]...), engine position, sync arrivals
    tags0 = (None, None) + tuple([None] * W) + tuple([None] * W)
    init = (tuple([0] * n_threads), tuple((0, BARS[b]) for b in BAR_NAMES), tags0, 0, tuple([0] * n_syncs))
    seen, todo = {init}, deque([init])
    deadlocks, violations = [], []

    def arrive(bars, name):
        i = BAR_NAMES.index(name)
        done, pend = bars[i]
        pend -= 1
        if pend == 0:
            done, pend = done + 1, BARS[name]
        return bars[:i] + ((done, pend),) + bars[i + 1:]

    while todo:
        st = todo.popleft()
        pcs, bars, tags, eng, syncs = st
        succ = []
        # the tensor-core engine retires the next issued MMA (in order, at any time after its issue)
        if eng < issued_before[pcs[0]]:
            _, _, src, dst, tag, commit = mmas[eng]
            if src[0] == "x":
                ok = all(tags[2 + w] == src[1] for w in range(W))
            else:
                ok = all(tags[2 + W + w] == src[1] for w in range(W))
            if not ok:
                violations.append(("operand", mmas[eng], tags))
            t2 = list(tags)
            t2[0 if dst == "D0" else 1] = tag
            succ.append((pcs, arrive(bars, commit), tuple(t2), eng + 1, syncs))
        for th in range(n_threads):
            if pcs[th] >= len(progs[th]):
                continue
            op = progs[th][pcs[th]]
            nxt = pcs[:th] + (pcs[th] + 1,) + pcs[th + 1:]
            if op[0] == "wait":
                done, _ = bars[BAR_NAMES.index(op[1])]
                if (done & 1) != op[2]:
                    succ.append((nxt, bars, tags, eng, syncs))
            elif op[0] == "sync":                  # __syncthreads: register the arrival (bit th), pass when all have
                k, full = op[1], (1 << n_threads) - 1
                if not (syncs[k] >> th) & 1:
                    succ.append((pcs, bars, tags, eng, syncs[:k] + (syncs[k] | (1 << th),) + syncs[k + 1:]))
                elif syncs[k] == full:
                    succ.append((nxt, bars, tags, eng, syncs))
            elif op[0] == "arrive":
                succ.append((nxt, arrive(bars, op[1]), tags, eng, syncs))
            elif op[0] == "issue":
                succ.append((nxt, bars, tags, eng, syncs))
            elif op[0] == "read":
                if tags[0 if op[1] == "D0" else 1] != op[2]:
                    violations.append(("accumulator", th, op, tags))
                succ.append((nxt, bars, tags, eng, syncs))
            elif op[0] == "writeA":
                t2 = list(tags)
                t2[2 + W + op[1]] = op[2]
                succ.append((nxt, bars, tuple(t2), eng, syncs))
            elif op[0] == "loadx":
                t2 = list(tags)
                t2[2 + op[1]] = op[2]
                succ.append((nxt, bars, tuple(t2), eng, syncs))
        finished = all(pcs[t] >= len(progs[t]) for t in range(n_threads)) and eng == len(mmas)
        if not succ and not finished:
            deadlocks.append(st)
        for s in succ:
            if s not in seen:
                seen.add(s)
                todo.append(s)
        assert len(seen) < limit, "state space larger than expected"
    return len(seen), deadlocks, violations


@pytest.mark.parametrize("nh", [1, 3])
@pytest.mark.parametrize("proto", ["per_tile", "ovl_own"])
def test_protocol_is_free_of_deadlock_and_hazards(proto, nh):
    states, deadlocks, violations = explore(proto, n_tiles=3, nh=nh)
    assert states > 100
    assert not deadlocks, f"{proto}: {len(deadlocks)} deadlocked states, e.g. {deadlocks[0]}"
    assert not violations, f"{proto}: {violations[0]}"


@pytest.mark.parametrize("nh", [1, 3])
def test_shared_barrier_overlap_can_hang(nh):
    """the withdrawn variant: the output-layer commit and the next tile's layer-0 commit both land on `bar` while a
    slow warp still sits in front of its wait for the first of them -> it sees its own parity again and never wakes"""
    _, deadlocks, _ = explore("ovl_shared", n_tiles=3, nh=nh)
    assert deadlocks, "the model did not reproduce the double phase completion"
    pcs, bars, _, _, _ = deadlocks[0]
    progs = build_programs("ovl_shared", 3, nh)
    stuck = [progs[t][pcs[t]] for t in range(len(progs)) if pcs[t] < len(progs[t])]
    assert any(op[0] == "wait" and op[1] == "bar" for op in stuck)


def test_even_hidden_layer_count_is_rejected_for_a_reason():
    """the kernels' static_assert(!OVL || (NH & 1)): with an even number of hidden layers the output accumulator is D0,
    which the next tile's layer-0 MMA overwrites while a slow warp still reads it — the model finds exactly that"""
    _, deadlocks, violations = explore("ovl_own", n_tiles=3, nh=2)
    assert not deadlocks and violations and all(v[0] == "accumulator" for v in violations)
    assert not explore("per_tile", n_tiles=3, nh=2)[2]          # the per-tile barrier has no such restriction


def test_three_warps_four_tiles(monkeypatch):
    """a larger instance of the same model (3 epilogue warps, 4 tiles): same verdicts"""
    import test_mbarrier_protocol_cpu as M

    monkeypatch.setattr(M, "W", 3)
    monkeypatch.setattr(M, "BARS", {"bar": 1, "bar_first": 1, "bar_x": 3, "chunk": 3})
    for proto in ("per_tile", "ovl_own"):
        states, deadlocks, violations = M.explore(proto, n_tiles=4, nh=3)
        assert states > 1000 and not deadlocks and not violations, proto
    assert M.explore("ovl_shared", n_tiles=4, nh=3)[1]


# ------------------------------------------------------------------------------------------------------------------
# generic checker (named regions, multi-region MMAs, named barriers) + the two-slot tone-map forward (k_tonemap_fwd2)
# ------------------------------------------------------------------------------------------------------------------
def explore_generic(progs, bar_counts, limit=4_000_000):
    """progs[0] = issuer.  Ops:
         ("wait", bar, parity) ("arrive", bar)
         ("sync", key, n_participants)                         named / CTA barrier instance `key`
         ("read", region, expected_tag) ("write", ((region, tag), ...))
         ("issue", reads ((region, expected), ...), writes ((region, tag), ...), commit_bar)   issuer only: tcgen05.mma
                                                               + commit, retired IN ORDER by the tensor-core engine
         ("tma", bar, ((copy_id, writes), ...))                mbarrier.arrive.expect_tx on `bar` + one cp.async.bulk per
                                                               copy; copies complete in ANY order (complete_tx)
    An mbarrier is (completed phases, pending arrivals, pending transactions): the phase completes when both reach 0.
    Returns (states, deadlocks, violations)."""
    bar_names = sorted(bar_counts)

    def op_regions(op):
        if op[0] == "read":
            return [op[1]]
        if op[0] == "write":
            return [x[0] for x in op[1]]
        if op[0] == "issue":
            return [x[0] for x in op[1]] + [x[0] for x in op[2]]
        if op[0] == "tma":
            return [x[0] for c in op[2] for x in c[1]]
        return []

    regions = sorted({r for p in progs for op in p for r in op_regions(op)})
    sync_keys = sorted({op[1] for p in progs for op in p if op[0] == "sync"}, key=repr)
    mmas = [op for op in progs[0] if op[0] == "issue"]
    copies = {c[0]: (op[1], c[1]) for p in progs for op in p if op[0] == "tma" for c in op[2]}
    issued_before, n = [], 0
    for op in progs[0]:
        issued_before.append(n)
        n += op[0] == "issue"
    issued_before.append(n)
    init = (tuple([0] * len(progs)), tuple((0, bar_counts[b], 0) for b in bar_names), tuple([None] * len(regions)), 0,
            tuple([0] * len(sync_keys)), frozenset())
    seen, todo, deadlocks, violations = {init}, deque([init]), [], []

    def update(bars, name, arrivals=0, tx=0):
        i = bar_names.index(name)
        done, pend, ptx = bars[i]
        pend, ptx = pend - arrivals, ptx + tx
        if pend == 0 and ptx == 0:
            done, pend = done + 1, bar_counts[name]
        return bars[:i] + ((done, pend, ptx),) + bars[i + 1:]

    def put(tags, writes):
        t2 = list(tags)
        for r, tag in writes:
            t2[regions.index(r)] = tag
        return tuple(t2)

    while todo:
        st = todo.popleft()
        pcs, bars, tags, eng, syncs, flying = st
        succ = []
        if eng < issued_before[pcs[0]]:
            _, reads, writes, commit = mmas[eng]
            for r, exp in reads:
                if tags[regions.index(r)] != exp:
                    violations.append(("operand", eng, r, exp, tags[regions.index(r)]))
            b2 = update(bars, commit, arrivals=1) if commit is not None else bars      # (None: covered by a later commit)
            succ.append((pcs, b2, put(tags, writes), eng + 1, syncs, flying))
        for cid in flying:                                   # any copy in flight may land next
            bar, writes = copies[cid]
            succ.append((pcs, update(bars, bar, tx=-1), put(tags, writes), eng, syncs, flying - {cid}))
        for th in range(len(progs)):
            if pcs[th] >= len(progs[th]):
                continue
            op = progs[th][pcs[th]]
            nxt = pcs[:th] + (pcs[th] + 1,) + pcs[th + 1:]
            if op[0] == "wait":
                if (bars[bar_names.index(op[1])][0] & 1) != op[2]:
                    succ.append((nxt, bars, tags, eng, syncs, flying))
            elif op[0] == "sync":
                k = sync_keys.index(op[1])
                if not (syncs[k] >> th) & 1:
                    succ.append((pcs, bars, tags, eng, syncs[:k] + (syncs[k] | (1 << th),) + syncs[k + 1:], flying))
                elif bin(syncs[k]).count("1") == op[2]:
                    succ.append((nxt, bars, tags, eng, syncs, flying))
            elif op[0] == "arrive":
                succ.append((nxt, update(bars, op[1], arrivals=1), tags, eng, syncs, flying))
            elif op[0] == "issue":
                succ.append((nxt, bars, tags, eng, syncs, flying))
            elif op[0] == "tma":
                b2 = update(update(bars, op[1], tx=len(op[2])), op[1], arrivals=1)
                succ.append((nxt, b2, tags, eng, syncs, flying | {c[0] for c in op[2]}))
            elif op[0] == "read":
                if tags[regions.index(op[1])] != op[2]:
                    violations.append(("read", th, op, tags[regions.index(op[1])]))
                succ.append((nxt, bars, tags, eng, syncs, flying))
            elif op[0] == "write":
                succ.append((nxt, bars, put(tags, op[1]), eng, syncs, flying))
        done = all(pcs[t] >= len(progs[t]) for t in range(len(progs))) and eng == len(mmas) and not flying
        if not succ and not done:
            deadlocks.append(st)
        for s2 in succ:
            if s2 not in seen:
                seen.add(s2)
                todo.append(s2)
        assert len(seen) < limit, "state space larger than expected"
    return len(seen), deadlocks, violations


def tonemap_fwd2_programs(n_my: int, n_warps: int, named_barrier: bool = True):
    """k_tonemap_fwd2 (mlp_tc.cu): two tiles in flight in slots s = i & 1.  Regions per slot: x_s[w] (encoded tile, one
    part per warp), D_s (layer-0 accumulator; the bf16 A operand is written OVER it, part A_s[w] per warp — the first
    A write destroys D_s for every warp, hence the named barrier between the D_s reads and the A_s writes), O_s."""
    issuer = []

    def mma0(i):
        s = i & 1
        issuer.append(("issue", tuple((f"x{s}_{w}", i) for w in range(n_warps)),
                       ((f"D{s}", (i, "d")),) + tuple((f"A{s}_{w}", None) for w in range(n_warps)), f"mma0_{s}"))

    for i in range(min(2, n_my)):
        issuer.append(("wait", f"x_{i}", 0))
        mma0(i)
    for i in range(n_my):
        s = i & 1
        issuer.append(("wait", f"a_{s}", (i >> 1) & 1))
        issuer.append(("issue", tuple((f"A{s}_{w}", (i, "a")) for w in range(n_warps)), ((f"O{s}", (i, "o")),), f"out_{s}"))
        if i + 2 < n_my:
            issuer.append(("wait", f"x_{s}", ((i + 2) >> 1) & 1))
            mma0(i + 2)
    warps = []
    for w in range(n_warps):
        ops = []

        def load_x(i):
            ops.append(("write", ((f"x{i & 1}_{w}", i),)))
            ops.append(("arrive", f"x_{i & 1}"))

        def out_epilogue(i):
            ops.append(("wait", f"out_{i & 1}", (i >> 1) & 1))
            ops.append(("read", f"O{i & 1}", (i, "o")))

        for i in range(min(2, n_my)):
            load_x(i)
        for i in range(n_my):
            s = i & 1
            ops.append(("wait", f"mma0_{s}", (i >> 1) & 1))
            if i + 2 < n_my:
                load_x(i + 2)
            ops.append(("read", f"D{s}", (i, "d")))
            if named_barrier:
                ops.append(("sync", ("named", i), n_warps))
            ops.append(("write", ((f"A{s}_{w}", (i, "a")), (f"D{s}", ("overwritten by A", i)))))
            ops.append(("arrive", f"a_{s}"))
            if i >= 1:
                out_epilogue(i - 1)
        if n_my >= 1:
            out_epilogue(n_my - 1)
        warps.append(ops)
    bars = {}
    for s in range(2):
        bars.update({f"x_{s}": n_warps, f"mma0_{s}": 1, f"a_{s}": n_warps, f"out_{s}": 1})
    return [issuer] + warps, bars


@pytest.mark.parametrize("n_my", [1, 2, 3, 5])
def test_tonemap_two_slot_forward_protocol(n_my):
    progs, bars = tonemap_fwd2_programs(n_my, n_warps=2)
    states, deadlocks, violations = explore_generic(progs, bars)
    assert states > 20 and not deadlocks and not violations, (deadlocks[:1], violations[:1])


def test_tonemap_two_slot_forward_needs_its_named_barrier():
    """without `bar.sync 1` between the D_s reads and the A_s writes a fast warp's A operand lands on accumulator
    columns a slow warp has not read yet: the model must see it (sensitivity check of the checker itself)"""
    progs, bars = tonemap_fwd2_programs(3, n_warps=2, named_barrier=False)
    _, deadlocks, violations = explore_generic(progs, bars)
    assert not deadlocks and any(v[0] == "read" for v in violations)


def wgrad_programs(n: int, boundary=(), wait_empty: bool = True):
    """k_mlp_wgrad_tc (mlp_tc.cu): split-K weight-gradient GEMM, two shared-memory stages filled by cp.async.bulk
    (operand tiles A_st, B_st; completion on full[st] via expect_tx) and drained by tcgen05.mma (commit on empty[st]).
    Thread 0 = the driver (producer + MMA issuer), thread 1 = the other threads of the CTA.  `boundary`: tiles whose
    out-of-range rows are zeroed in shared memory by all threads before the MMAs and whose "ones" rows are restored after
    them.  `wait_empty=False` drops the driver's wait for the MMAs of tile it - 1 before it refills their stage."""
    driver, other = [], []

    def load(tile, st):
        driver.append(("tma", f"full{st}", ((("A", tile), ((f"A{st}", tile),)), (("B", tile), ((f"B{st}", tile),)))))

    load(0, 0)
    for it in range(n):
        st, par = it & 1, (it >> 1) & 1
        if it + 1 < n:
            if it >= 1 and wait_empty:
                driver.append(("wait", f"empty{st ^ 1}", ((it - 1) >> 1) & 1))
            load(it + 1, st ^ 1)
        if it in boundary:
            for t in (driver, other):
                t.append(("wait", f"full{st}", par))
                t.append(("read", f"A{st}", it))                 # the rows being zeroed are this tile's
                t.append(("sync", ("zeroed", it), 2))
        driver.append(("wait", f"full{st}", par))
        driver.append(("issue", ((f"A{st}", it), (f"B{st}", it)), (("acc", it),), f"empty{st}"))
        for t in (driver, other):
            t.append(("sync", ("iter", it), 2))
        if it in boundary:
            for t in (driver, other):
                t.append(("wait", f"empty{st}", par))
                t.append(("read", f"B{st}", it))                 # restoring the "ones" rows: the stage is still this tile's
                t.append(("sync", ("restored", it), 2))
    for t in (driver, other):
        t.append(("wait", f"empty{(n - 1) & 1}", ((n - 1) >> 1) & 1))
        t.append(("read", "acc", n - 1))
    return [driver, other], {"full0": 1, "full1": 1, "empty0": 1, "empty1": 1}


@pytest.mark.parametrize("n,boundary", [(1, ()), (2, ()), (5, ()), (4, (0, 3)), (1, (0,)), (5, (0, 1, 2, 3, 4))])
def test_weight_gradient_pipeline_protocol(n, boundary):
    progs, bars = wgrad_programs(n, boundary)
    states, deadlocks, violations = explore_generic(progs, bars)
    assert states > 10 and not deadlocks and not violations, (deadlocks[:1], violations[:1])


def test_weight_gradient_pipeline_needs_the_empty_wait():
    """sensitivity: refilling a stage without waiting for the MMAs that still read it is seen by the model"""
    progs, bars = wgrad_programs(5, (), wait_empty=False)
    _, _, violations = explore_generic(progs, bars)
    assert any(v[0] == "operand" for v in violations)


def tonemap_bwd_programs(n_tiles: int, n_warps: int, wait_bar_w: bool = True):
    """k_tonemap_bwd_fused (mlp_tc.cu): per tile T0 (encode X, dZ_out) | MMA X W0 -> D | E1 (H -> Hs) | MMA dZ_out Wo^T -> D,
    dWo^T += Hs^T dZ_out | E2 (dZ0 -> A, Zs) | MMA A W0 -> S, dW0 += Zs^T [X|1] (commit bar_w) | E3; a __syncthreads
    after T0, E1, E2; the weight-gradient MMAs of tile t still read X / Hs / Zs / dZ_out while tile t + 1 begins, so T0
    of the next tile first waits for bar_w."""
    nthreads = n_warps + 1
    issuer, warps = [], [[] for _ in range(n_warps)]
    every = lambda name, t: tuple((f"{name}_{w}", t) for w in range(n_warps))
    for t in range(n_tiles):
        for w, ops in enumerate(warps):
            if t > 0 and wait_bar_w:
                ops.append(("wait", "bar_w", (t - 1) & 1))
            ops.append(("write", ((f"X_{w}", t), (f"DZO_{w}", t))))
            ops.append(("sync", ("T0", t), nthreads))
        issuer.append(("sync", ("T0", t), nthreads))
        issuer.append(("issue", every("X", t), (("D", (t, "z0")),), "bar"))
        for w, ops in enumerate(warps):
            ops.append(("wait", "bar", (3 * t) & 1))
            ops.append(("read", "D", (t, "z0")))
            ops.append(("write", ((f"HS_{w}", t),)))
            ops.append(("sync", ("E1", t), nthreads))
        issuer.append(("sync", ("E1", t), nthreads))
        issuer.append(("issue", every("DZO", t), (("D", (t, "dh")),), "bar"))
        issuer.append(("issue", every("HS", t) + every("DZO", t), (("GO", t),), None))
        for w, ops in enumerate(warps):
            ops.append(("wait", "bar", (3 * t + 1) & 1))
            ops.append(("read", "D", (t, "dh")))
            ops.append(("write", ((f"A_{w}", t), (f"ZS_{w}", t))))
            ops.append(("sync", ("E2", t), nthreads))
        issuer.append(("sync", ("E2", t), nthreads))
        issuer.append(("issue", every("A", t), (("S", t),), "bar"))
        issuer.append(("issue", every("ZS", t) + every("X", t), (("G0", t),), "bar_w"))
        for w, ops in enumerate(warps):
            ops.append(("wait", "bar", (3 * t + 2) & 1))
            ops.append(("read", "S", t))
    for ops in warps:
        ops.append(("wait", "bar_w", (n_tiles - 1) & 1))
        ops.append(("read", "G0", n_tiles - 1))
        ops.append(("read", "GO", n_tiles - 1))
    return [issuer] + warps, {"bar": 1, "bar_w": 1}


@pytest.mark.parametrize("n_tiles", [1, 2, 4])
def test_fused_tonemap_backward_protocol(n_tiles):
    progs, bars = tonemap_bwd_programs(n_tiles, n_warps=2)
    states, deadlocks, violations = explore_generic(progs, bars)
    assert states > 20 and not deadlocks and not violations, (deadlocks[:1], violations[:1])


def test_fused_tonemap_backward_needs_the_bar_w_wait():
    progs, bars = tonemap_bwd_programs(3, n_warps=2, wait_bar_w=False)
    _, _, violations = explore_generic(progs, bars)
    assert any(v[0] == "operand" for v in violations)


def chain_programs_chunked(ovl: bool, n_tiles: int, nh: int, n_warps: int, chunks: int = 3, dgrad: bool = False):
    """the forward chain again, with its chunk pipeline spelt out (mlp_tc.cu `chunk_ready`): the A operand is produced in
    `chunks` column chunks per warp, each with its own mbarrier (count = warps, one shared parity); the issuer fires the
    K-steps of chunk cc as soon as chunk cc is complete and commits once after the last chunk.
    `dgrad`: the data-gradient chain (k_mlp_dgrad_tc) — the first MMA's shared-memory operand is the dZ_out tile, which
    each tile's warps write at the top of the tile (per-tile barrier) or, with the overlap, for the NEXT tile after
    the chain epilogues and before the arrival on bar_x (`make_dz(tile + gridDim.x)`)"""
    d = lambda i: f"D{i & 1}"
    issuer, cphase, xphase = [], 0, 0

    def layer0(t):
        issuer.append(("issue", tuple((f"x_{w}", t) for w in range(n_warps)), ((d(0), (t, 0)),), "bar_first" if ovl else "bar"))

    for t in range(n_tiles):
        more = t + 1 < n_tiles
        if not ovl or t == 0:
            issuer.append(("sync", ("tile", t), n_warps + 1))
            layer0(t)
        for l in range(nh):
            for cc in range(chunks):
                issuer.append(("wait", f"chunk{cc}", cphase))
                last = cc == chunks - 1
                issuer.append(("issue", tuple((f"A_{w}_{cc}", (t, l)) for w in range(n_warps)),
                               ((d(l + 1), (t, l + 1) if last else (t, l + 1, "partial", cc)),), "bar" if last else None))
            cphase ^= 1
        if ovl and more:
            issuer.append(("wait", "bar_x", xphase))
            xphase ^= 1
            layer0(t + 1)
    warps = []
    for w in range(n_warps):
        ops, phase, fphase = ([] if dgrad else [("write", ((f"x_{w}", 0),))]), 0, 0
        for t in range(n_tiles):
            more = t + 1 < n_tiles
            if dgrad and (not ovl or t == 0):
                ops.append(("write", ((f"x_{w}", t),)))          # dZ_out tile of this tile
            if not ovl or t == 0:
                ops.append(("sync", ("tile", t), n_warps + 1))
            for l in range(nh):
                if l == 0 and ovl:
                    ops.append(("wait", "bar_first", fphase))
                    fphase ^= 1
                else:
                    ops.append(("wait", "bar", phase))
                    phase ^= 1
                if l == 0 and more and not dgrad:
                    ops.append(("write", ((f"x_{w}", t + 1),)))
                ops.append(("read", d(l), (t, l)))              # all three chunks are loaded from TMEM up front
                for cc in range(chunks):
                    ops.append(("write", ((f"A_{w}_{cc}", (t, l)),)))
                    ops.append(("arrive", f"chunk{cc}"))
            if ovl and more:
                if dgrad:
                    ops.append(("write", ((f"x_{w}", t + 1),)))
                ops.append(("arrive", "bar_x"))
            ops.append(("wait", "bar", phase))
            phase ^= 1
            ops.append(("read", d(nh), (t, nh)))
        warps.append(ops)
    bars = {"bar": 1, "bar_first": 1, "bar_x": n_warps}
    bars.update({f"chunk{cc}": n_warps for cc in range(chunks)})
    return [issuer] + warps, bars


@pytest.mark.parametrize("dgrad", [False, True])
@pytest.mark.parametrize("ovl", [False, True])
def test_chunk_pipelined_chain_protocol(ovl, dgrad):
    progs, bars = chain_programs_chunked(ovl, n_tiles=3, nh=3, n_warps=2, dgrad=dgrad)
    states, deadlocks, violations = explore_generic(progs, bars)
    assert states > 1000 and not deadlocks and not violations, (deadlocks[:1], violations[:1])
