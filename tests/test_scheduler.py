"""Work-unit lists of the grouped GEMM (host side, no GPU): whatever the tiler does -- LPT order, split-K chosen by the
simulated schedule, forced cuts for staggered output, row-slab partitions -- every output element must be produced by
exactly one tile, every tile's k loop covered exactly once by its units in slot order, and the launch order must be
costliest first.  (The kernels that consume these lists are covered by the GPU parity tests.)"""
import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import _lib, workloads as wl


def heff_matches(ix, dtype):
    rng = np.random.default_rng(2)
    ti = wl.heff_tensor_indexes(ix)
    div = (0,) * ti["psi"][0].kind.nvals
    shells = {n: tk.BlockSparseTensor(i, dtype).random(div, rng) for n, i in ti.items()}
    out = []
    for lhs, rhs, axes, res in wl.HEFF_STEPS:
        m = tk.Match(shells[lhs], shells[rhs], axes)
        shells[res] = m.result_shell(dtype)
        out.append(m)
    return out


def check_units(plan, match, dtype, stream_k=False):
    units, (bm, bn, bk) = plan.units()
    if not units:
        return 0
    tasks = match.tasks(sorted_by_c=True)
    stages, shape = {}, {}
    for t in tasks:                               # k loop of an output block = its pairs' k loops back to back, in stages of bk
        stages[t.c_ord] = stages.get(t.c_ord, 0) + -(-int(t.k) // bk)
        shape[t.c_ord] = (int(t.m), int(t.n))
    # groups are numbered like the output blocks that have DMMA work; map through (rows, cols) coverage instead of ids
    by_tile = {}
    for u in units:
        by_tile.setdefault((u.group, u.tm, u.tn), []).append(u)
    cover = {}
    for (g, tm, tn), us in by_tile.items():
        us = sorted(us, key=lambda u: u.split)
        assert [u.split for u in us] == list(range(len(us))) and all(u.nsplit == len(us) for u in us)
        assert us[0].s_begin == 0
        for a, b in zip(us, us[1:]):
            assert a.s_end == b.s_begin and a.s_end > a.s_begin      # contiguous, non-empty, slot order == k order
        assert us[-1].s_end > us[-1].s_begin
        assert 1 <= us[0].rows <= bm and 1 <= us[0].cols <= bn
        cover.setdefault(g, {"stages": us[-1].s_end, "tiles": set(), "elems": 0})
        assert cover[g]["stages"] == us[-1].s_end                    # every tile of a block walks the same k loop
        assert (tm, tn) not in cover[g]["tiles"]
        cover[g]["tiles"].add((tm, tn))
        cover[g]["elems"] += us[0].rows * us[0].cols
    if stream_k:
        return sum(c["elems"] for c in cover.values())
    # launch order: costliest first (k-loop length x issued share of the tile), ties keep generation order
    def cost(u):
        mt, nt = -(-u.rows // 8), -(-(-(-u.cols // 8)) // 4)
        return (u.s_end - u.s_begin) * mt * nt
    costs = [cost(u) for u in units]
    assert all(a >= b for a, b in zip(costs, costs[1:]))
    return sum(c["elems"] for c in cover.values())


def check_items(plan):
    """Narrow-pair work items: the items of a block are consecutive row ranges starting at row 0, none empty, none longer than
    the kernel's item size; returns the output elements they cover."""
    by_group = {}
    for g, row0, rows, n in plan.items():
        assert rows >= 1 and 1 <= n <= 8
        by_group.setdefault(g, []).append((row0, rows, n))
    elems = 0
    for g, its in by_group.items():
        its.sort()
        assert its[0][0] == 0
        for (r0, nr, n), (r1, _, n1) in zip(its, its[1:]):
            assert r0 + nr == r1 and n == n1                         # contiguous, no overlap
        per = {nr for _, nr, _ in its[:-1]}
        assert len(per) <= 1 and all(its[-1][1] <= x for x in per)   # equal items, a shorter tail
        elems += sum(nr * n for _, nr, n in its)
    return elems


@pytest.mark.parametrize("per_slot", ["1", "2", "8"])
def test_narrow_pair_items_tile_their_blocks(per_slot, monkeypatch):
    """The MPO-application steps of the chain at two sizes, with the item-size tuning switch at several settings."""
    monkeypatch.setenv("QLB200_SKINNY_ITEMS_PER_SLOT", per_slot)
    seen = 0
    for D in (300, 1500):
        ms = heff_matches(wl.u1_heisenberg_indexes(D), np.complex128)
        for m in ms:
            plan = tk.ContractionPlan(None, m, np.complex128, _lib.PLAN_DETERMINISTIC)
            st = plan.stats()
            n = check_items(plan)
            if st.ntile_dmma == 0:
                assert n == m.c_elems                                # a pure narrow-pair step: the items are all of C
                seen += 1
            plan.close(); m.close()
    assert seen == 4                                                 # steps 2 and 3 of both chains


@pytest.mark.parametrize("dtype", [np.complex128, np.float64])
@pytest.mark.parametrize("flags", [0, _lib.PLAN_STAGGER_OUTPUT, _lib.PLAN_NO_SPLIT_K, _lib.PLAN_CPLX_4M, _lib.PLAN_NO_SKINNY | _lib.PLAN_STAGGER_OUTPUT,
                                   _lib.PLAN_STREAM_K, _lib.PLAN_STREAM_K | _lib.PLAN_CPLX_4M | _lib.PLAN_NO_SKINNY])
@pytest.mark.parametrize("workload", ["u1_300", "u1_1500", "hubbard_200"])
def test_units_cover_every_tile_once(workload, flags, dtype):
    ix = {"u1_300": lambda: wl.u1_heisenberg_indexes(300), "u1_1500": lambda: wl.u1_heisenberg_indexes(1500),
          "hubbard_200": lambda: wl.hubbard_indexes(200)}[workload]()
    for m in heff_matches(ix, dtype):
        plan = tk.ContractionPlan(None, m, dtype, _lib.PLAN_DETERMINISTIC | flags)
        st = plan.stats()
        covered = check_units(plan, m, dtype, stream_k=bool(flags & _lib.PLAN_STREAM_K))
        if st.nrow_skinny == 0 and st.ntile_dmma:
            assert covered == m.c_elems                              # all of C comes from DMMA tiles
        assert covered + check_items(plan) == m.c_elems              # DMMA tiles + narrow-pair items = every element of C, once
        plan.close(); m.close()


@pytest.mark.parametrize("world", [2, 8])
def test_partitioned_units_tile_the_rank_slabs(world):
    """Row-slab partitions (multi-GPU): the ranks' tiles cover disjoint rows that add up to the whole result."""
    (m, *rest) = heff_matches(wl.u1_heisenberg_indexes(1200), np.complex128)
    total = 0
    for rank in range(world):
        plan = tk.ContractionPlan(None, m, np.complex128, _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SKINNY | _lib.PLAN_STAGGER_OUTPUT)
        plan.partition(world, rank)
        total += check_units(plan, m, np.complex128)
        off, ln = plan.c_ranges()
        assert sum(int(x) for x in ln) > 0 or world > 4
        plan.close()
    assert total == m.c_elems
    for x in [m] + rest:
        x.close()


def test_stagger_cuts_long_loops_and_keeps_a_tile_together():
    """QLB200_PLAN_STAGGER_OUTPUT: long k loops are cut into about four units, queued back to back."""
    (*rest, m) = heff_matches(wl.u1_heisenberg_indexes(2048), np.complex128)      # the last step: k runs over (wb, vb)
    plain = tk.ContractionPlan(None, m, np.complex128, _lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_SPLIT_K)
    stag = tk.ContractionPlan(None, m, np.complex128, _lib.PLAN_DETERMINISTIC | _lib.PLAN_STAGGER_OUTPUT)
    up, _ = plain.units()
    us, _ = stag.units()
    longest = max(u.s_end for u in up)
    assert all(u.nsplit == 1 for u in up)
    assert longest > 32
    assert max(u.nsplit for u in us) >= 4 and max(u.s_end - u.s_begin for u in us) <= max(8, -(-longest // 4)) + 1
    pos = {}
    for i, u in enumerate(us):
        pos.setdefault((u.group, u.tm, u.tn), []).append(i)
    for idxs in pos.values():                                       # equal-cost units of one tile stay adjacent (stable order)
        lens = {us[i].s_end - us[i].s_begin for i in idxs}
        if len(lens) == 1:
            assert idxs == list(range(idxs[0], idxs[0] + len(idxs)))
    plain.close(); stag.close()
    for x in [m] + rest:
        x.close()


def unit_cost(u):
    mt, nt = -(-u.rows // 8), -(-(-(-u.cols // 8)) // 4)
    return (u.s_end - u.s_begin) * mt * nt


@pytest.mark.parametrize("world,rank", [(1, 0), (8, 1), (8, 4)])
def test_stream_k_segments_are_balanced(world, rank):
    """QLB200_PLAN_STREAM_K: one contiguous unit range per resident CTA, in order, covering every unit; the costliest
    segment stays within a few percent of the mean, where the default LPT list of whole tiles cannot when there are
    about as many equal tiles as CTA slots (the 8-GPU shards of the headline workload)."""
    from tensortoolkit_b200.sharding import shard_chain
    rng = np.random.default_rng(5)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(4096))
    t = {n: tk.BlockSparseTensor(i, np.complex128) for n, i in ti.items()}
    for n in t:                                            # structure only: block lists without touching 100 MB of data
        t[n].set_blocks(t[n].div_blocks((0,)))
    if world > 1:
        t, _ = shard_chain(t, wl.HEFF_STEPS, "lenv", 2, world, rank, np.complex128)
    shells = dict(t)
    for step, (lhs, rhs, axes, res) in enumerate(wl.HEFF_STEPS):
        m = tk.Match(shells[lhs], shells[rhs], axes)
        shells[res] = m.result_shell(np.complex128)
        if step in (0, 3):
            sk = tk.ContractionPlan(None, m, np.complex128, _lib.PLAN_DETERMINISTIC | _lib.PLAN_STREAM_K)
            units, _ = sk.units()
            seg = sk.segments()
            assert seg[0] == 0 and seg[-1] == len(units) and all(a < b for a, b in zip(seg, seg[1:]))
            assert len(seg) - 1 <= 296
            costs = [sum(unit_cost(u) for u in units[a:b]) for a, b in zip(seg, seg[1:])]
            mean = sum(costs) / 296.0
            assert max(costs) <= 1.06 * mean, (step, max(costs) / mean)
            # at most two cut tiles per segment boundary: the fix-up traffic stays small
            assert sum(1 for u in units if u.nsplit > 1) <= 3 * len(costs)
            sk.close()
        m.close()


def test_ragged_b_blocks_are_read_as_views():
    """BASELINE configs[4]: B stored (n1, k, n2), contracted over k.  The k x (n1 n2) matrix is a strided view of the stored
    block (rows = n1 runs of n2 contiguous elements), which the GEMM producers read in place: the plan sends (almost) nothing
    through the permute kernel; QLB200_PLAN_NO_VIEW restores the permute pass (86 % of B)."""
    import numpy as np
    import tensortoolkit_b200 as tk
    from tensortoolkit_b200 import _lib, workloads as wl
    tb = wl.ragged_tables()
    mk = lambda fl: tk.RawPlan(None, np.float64, 3, [1, 2, 0], tb["a_shape"], tb["a_off"], 3, [1, 0, 2], tb["b_shape"], tb["b_off"],
                               tb["tasks"], tb["c_elems"], fl)
    view, noview = mk(_lib.PLAN_DETERMINISTIC), mk(_lib.PLAN_DETERMINISTIC | _lib.PLAN_NO_VIEW)
    sv, sn = view.stats(), noview.stats()
    assert sv.permute_elems_a == 0 and sn.permute_elems_a == 0
    assert sv.permute_elems_b < 1e-3 * tb["b_elems"]
    assert sn.permute_elems_b > 0.8 * tb["b_elems"]
    assert sv.ntile_dmma == sn.ntile_dmma and sv.flops == sn.flops
    # every block is either read in place or has a workspace slot
    for b in (0, 17, 4321):
        assert (view.operand_block(1, b) is None) or sv.permute_elems_b > 0
    view.close(); noview.close()
