"""Multi-GPU partition logic (host side): every rank computes the same cuts, the slabs of all
ranks tile the full result exactly, and the all-gather + unpack rebuilds it.  The 2-process case
runs over gloo on CPU with the numpy oracle standing in for the local contraction kernels."""
import os
import socket

import numpy as np
import pytest

import tensortoolkit_b200 as tk
from tensortoolkit_b200 import sharding as sh, workloads as wl


def make_tensors(D, dtype, seed):
    rng = np.random.default_rng(seed)
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(D))
    return {n: tk.BlockSparseTensor(idxs, dtype).random((0,), rng) for n, idxs in ti.items()}


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slabs_tile_the_result(world):
    ts = make_tensors(300, np.float64, 1)
    infos = [sh.shard_heff_tensors(ts, world, r)[1] for r in range(world)]
    cov = np.zeros(infos[0].full_elems, np.int32)
    for r in range(world):
        assert infos[r].sector_ranges == infos[0].sector_ranges      # identical cuts on every rank
        loc = 0
        for s in infos[0].slabs[r]:
            assert s.local_offset == loc
            loc += s.length
            cov[s.full_offset:s.full_offset + s.length] += 1
        assert loc == infos[0].local_elems[r]
    assert np.all(cov == 1)
    assert abs(sum(infos[0].cost_share) - 1.0) < 1e-9
    if world > 1:
        assert max(infos[0].cost_share) * world < 1.25       # balanced well beyond whole-sector LPT (~1.36 at 8)


@pytest.mark.parametrize("D,world", [(17, 8), (20, 8), (24, 8), (40, 8), (40, 5), (64, 8), (8, 4)])
def test_small_bonds_leave_idle_ranks_not_crashes(D, world):
    """world may exceed what the snapped cuts can separate: an empty share is legal (the rank idles), every non-empty
    share still matches through the whole chain, and the slabs of all ranks tile the result exactly."""
    ts = make_tensors(D, np.float64, 4)
    cov = None
    n_idle = 0
    for r in range(world):
        mine, info = sh.shard_heff_tensors(ts, world, r)
        if cov is None:
            cov = np.zeros(info.full_elems, np.int32)
        if info.local_elems[r] == 0:
            n_idle += 1
            assert info.slabs[r] == []
            continue
        shells = dict(mine)
        for lhs, rhs, axes, out in wl.HEFF_STEPS:              # the host matcher accepts every restricted operand
            m = tk.Match(shells[lhs], shells[rhs], axes)
            shells[out] = m.result_shell(np.float64)
            m.close()
        assert shells["out"].data.size == info.local_elems[r]
        for s in info.slabs[r]:
            cov[s.full_offset:s.full_offset + s.length] += 1
    assert np.all(cov == 1)
    assert n_idle < world


def test_partition_feedback_moves_rows_from_slow_ranks():
    """sharding.reweigh_pieces: after the step the modelled share of every rank equals its measured share (damp = 1) or lies
    between the old and the measured one (damp < 1); the re-cut line still tiles every sector exactly once."""
    degs = [40, 200, 680, 200, 40]
    cost = np.array([1.0, 2.0, 3.0, 2.0, 1.0])          # per-row cost of each sector
    world = 4
    pieces = sh.line_pieces(cost, degs)
    cuts = sh.cut_line(pieces, len(degs), degs, world, snap=8)
    times = [1.0, 1.3, 0.9, 1.1]                        # rank 1 is slow, rank 2 fast

    def model(pcs, ranges):
        out = []
        for r in range(world):
            w = 0.0
            for s, lo, hi, wt in pcs:
                a, b = max(lo, ranges[r][s][0]), min(hi, ranges[r][s][1])
                w += max(0, b - a) * wt
            out.append(w)
        return np.array(out) / sum(out)

    before = model(pieces, cuts)
    want = np.array(times) / sum(times)                  # measured shares
    full = model(sh.reweigh_pieces(pieces, cuts, times, damp=1.0), cuts)
    assert np.allclose(full, want, rtol=1e-12)
    half = model(sh.reweigh_pieces(pieces, cuts, times, damp=0.5), cuts)
    for r in range(world):
        lo, hi = sorted((before[r], want[r]))
        assert lo - 1e-12 <= half[r] <= hi + 1e-12
    # the slow rank's rows got heavier: an equal cut of the new line gives it fewer rows
    new_cuts = sh.cut_line(sh.reweigh_pieces(pieces, cuts, times, damp=0.5), len(degs), degs, world, snap=8)
    rows = lambda c, r: sum(hi - lo for lo, hi in c[r])
    assert rows(new_cuts, 1) < rows(cuts, 1) and rows(new_cuts, 2) > rows(cuts, 2)
    for s, d in enumerate(degs):                         # every sector still tiled exactly once
        segs = sorted((c[s] for c in new_cuts if c[s][1] > c[s][0]))
        assert segs[0][0] == 0 and segs[-1][1] == d and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))


def test_tune_partition_keeps_the_best_measured_cut():
    """sharding.tune_partition on a synthetic machine whose rank 1 is 30 % slow and whose time has a per-tile step (64 rows):
    the kept cut is never worse than the flop-balanced one, every abandoned chain is closed, and the loop is deterministic."""
    import types
    degs = [40, 200, 680, 200, 40]
    cost = np.array([1.0, 2.0, 3.0, 2.0, 1.0])
    world = 4
    slow = [1.0, 1.3, 1.0, 1.0]
    closed = []

    def build(pieces):
        cuts = sh.cut_line(pieces, len(degs), degs, world, snap=8)
        c = types.SimpleNamespace(info=types.SimpleNamespace(pieces=pieces, sector_ranges=cuts))
        c.close = lambda c=c: closed.append(id(c))
        return c

    def measure(c):
        out = []
        for r in range(world):
            t = 0.0
            for s, (lo, hi) in enumerate(c.info.sector_ranges[r]):
                t += -(-(hi - lo) // 64) * 64 * cost[s]      # whole tiles of 64 rows
            out.append(t * slow[r])
        return out

    first = build(sh.line_pieces(cost, degs))
    t0 = max(measure(first))
    chain, kept, log = sh.tune_partition(first, build, measure, rounds=4, damp=0.5)
    assert len(log) == 5 and max(measure(chain)) == min(max(t) for t in log) <= t0
    assert max(log[kept]) == max(measure(chain))
    assert max(measure(chain)) < t0                        # the slow rank did get fewer rows
    assert len(closed) == (4 if kept == 4 else 5) and id(chain) not in closed
    chain2, kept2, log2 = sh.tune_partition(build(sh.line_pieces(cost, degs)), build, measure, rounds=4, damp=0.5)
    assert kept2 == kept and log2 == log


def test_c_partitioner_equals_the_python_restatement():
    """qlb200_shard_cut_line / qlb200_shard_reweigh (the product) against their Python restatement (tests/util.py) on random
    lines: identical cuts (integers) and identical pieces (weights to the last bit: same operations in the same order)."""
    from tests import util
    rng = np.random.default_rng(5)
    for case in range(200):
        nsct = int(rng.integers(1, 20))
        degs = [int(x) for x in rng.integers(1, 700, nsct)]
        cost = rng.random(nsct) * rng.integers(0, 2, nsct).clip(0, 1) if case % 7 == 0 else rng.random(nsct) + 0.01
        world = int(rng.integers(1, 9))
        snap = int(rng.choice([1, 4, 8, 16]))
        pieces = sh.line_pieces(cost * np.array(degs), degs)
        cuts = sh.cut_line(pieces, nsct, degs, world, snap)
        assert cuts == util.cut_line_py(pieces, nsct, degs, world, snap), case
        times = list(rng.random(world) + 0.5)
        for damp in (1.0, 0.5):
            got = sh.reweigh_pieces(pieces, cuts, times, damp)
            want = util.reweigh_pieces_py(pieces, cuts, times, damp)
            assert [(a, b, c) for a, b, c, _ in got] == [(a, b, c) for a, b, c, _ in want], case
            assert np.array_equal([w for *_, w in got], [w for *_, w in want]) or np.allclose([w for *_, w in got], [w for *_, w in want], rtol=1e-15, atol=0), case
            cuts2 = sh.cut_line(got, nsct, degs, world, snap)
            assert cuts2 == util.cut_line_py(want, nsct, degs, world, snap), case


def test_sector_costs_attribute_every_flop_once():
    """qlb200_shard_sector_flops over the four steps of the chain: the line's total weight is the chain's flop count
    (qlb200_estimate_cost), and it equals a task-by-task restatement."""
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(120))
    rng = np.random.default_rng(3)
    t = {n: tk.BlockSparseTensor(ix, np.float64).random((0, 0), rng) for n, ix in ti.items()}
    cost, shells, out_axis = sh.sector_costs(t, wl.HEFF_STEPS, "lenv", 2, np.float64)
    want = np.zeros_like(cost)
    total = 0.0
    cur = dict(t)
    track, _ = sh._track_axis(wl.HEFF_STEPS, {k: v.rank for k, v in t.items()}, "lenv", 2)
    for (lhs, rhs, axes, res), p in zip(wl.HEFF_STEPS, track):
        m = tk.Match(cur[lhs], cur[rhs], axes)
        for task in m.tasks():
            want[int(cur[lhs].blk_coors[task.a_ord, p])] += 2.0 * task.m * task.k * task.n
        total += m.cost(np.float64).flops if hasattr(m, "cost") else sum(2.0 * x.m * x.k * x.n for x in m.tasks())
        cur[res] = m.result_shell(np.float64)
        m.close()
    assert np.allclose(cost, want, rtol=1e-14) and np.isclose(cost.sum(), total, rtol=1e-12)


def test_c_row_slab_equals_the_numpy_restatement():
    """qlb200_shard_restrict (structure + copy list, applied by sharding.restrict_tensor) against block-by-block numpy slicing:
    every axis of a rank-4 fermionic tensor, random row ranges incl. empty and full sectors."""
    from tests import util
    rng = np.random.default_rng(11)
    ti = wl.heff_tensor_indexes(wl.hubbard_indexes(60))
    for name, div in (("psi", (0, 0)), ("lenv", (0, 0))):
        t = tk.BlockSparseTensor(ti[name], np.complex128).random(div, rng)
        for axis in range(t.rank):
            for trial in range(4):
                ranges = []
                for sct in t.indexes[axis].sectors:
                    d = int(sct.dgnc)
                    kind = rng.integers(0, 4)
                    lo, hi = (0, d) if kind == 0 else (0, 0) if kind == 1 else sorted(int(x) for x in rng.integers(0, d + 1, 2))
                    ranges.append((lo, hi))
                got, want = sh.restrict_tensor(t, axis, ranges), util.restrict_tensor_py(t, axis, ranges)
                assert got.same_structure(want) and got.indexes == want.indexes
                assert np.array_equal(got.data, want.data)


def test_cpp_client_shards_step_one_like_the_python_face(tmp_path):
    """tests/cpp/shard_host.cc: a C++ program using only include/qlb200.h reads the shells of lenv and psi, matches, attributes
    the flops, cuts the row line, builds every rank's slab and matches it -- same rows, slab sizes, task counts and result
    sizes as tensortoolkit_b200/sharding.py."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ti = wl.heff_tensor_indexes(wl.u1_heisenberg_indexes(700))
    rng = np.random.default_rng(4)
    t = {n: tk.BlockSparseTensor(ti[n], np.complex128).random((0,), rng) for n in ("lenv", "psi")}
    path = tmp_path / "shells.txt"
    with open(path, "w") as f:
        for name in ("lenv", "psi"):
            x = t[name]
            f.write(f"{x.rank} {x.nblk}\n")
            for ix in x.indexes:
                f.write(f"{ix.nsct} {int(ix.dir)}\n")
            f.write(" ".join(str(int(d)) for ix in x.indexes for d in ix.degs()) + "\n")
            f.write(" ".join(str(int(c)) for c in x.blk_coors.reshape(-1)) + "\n")
    exe = tmp_path / "shard_host"
    lib_dir = os.path.join(root, "tensortoolkit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "shard_host.cc"), "-o", str(exe), "-L", lib_dir, "-lqlb200", "-Wl,-rpath," + lib_dir])
    world = 4
    out = subprocess.run([str(exe), str(path), str(world)], capture_output=True, text=True, check=True).stdout.splitlines()
    steps = [("lenv", "psi", ([0], [0]), "t1")]
    m = tk.Match(t["lenv"], t["psi"], ([0], [0]))
    assert out[0] == f"whole: tasks {len(m.tasks())} c_elems {m.c_elems}"
    m.close()
    cost, _, _ = sh.sector_costs(t, steps, "lenv", 2, np.complex128)
    degs = [int(d) for d in t["lenv"].indexes[2].degs()]
    cuts = sh.row_line_cuts(cost, degs, world, snap=8)
    for r in range(world):
        slab = sh.restrict_tensor(t["lenv"], 2, cuts[r])
        rows = sum(hi - lo for lo, hi in cuts[r])
        copies = len(sh.slab_layout(t["lenv"], 2, cuts[r])[4][0])
        ms = tk.Match(slab, t["psi"], ([0], [0])) if slab.nblk else None
        want = f"rank {r}: rows {rows} slab_elems {slab.data.size if slab.nblk else 0} copies {copies} tasks {len(ms.tasks()) if ms else 0} c_elems {ms.c_elems if ms else 0}"
        assert out[1 + r] == want
        if ms:
            ms.close()


def test_cpp_adapter_row_slab_and_cuts_on_reference_tensors(ref):
    """include/qlten_b200/sharding.h instantiated on the reference's own QLTensor types (oracle/_ref/libqladapter.so, host only):
    qlten::b200::RowSlab gives the tensor sharding.restrict_tensor gives -- same indexes (checked by the reference's own
    Index ==), same blocks, same raw data -- for bosonic and fermionic tensors, and SectorFlops + CutRowLine give sharding.py's
    cuts."""
    rng = np.random.default_rng(21)
    for ixf, div, dtype in ((wl.u1_heisenberg_indexes, (0,), np.complex128), (wl.hubbard_indexes, (0, 0), np.float64)):
        ti = wl.heff_tensor_indexes(ixf(90))
        ref.set_seed(5)
        lenv = ref.RefTensor.new(ti["lenv"], dtype).random(div)
        psi = ref.RefTensor.new(ti["psi"], dtype).random(div)
        tl, tp = lenv.to_bst(), psi.to_bst()
        steps = [("lenv", "psi", ([0], [0]), "t1")]
        cost, _, _ = sh.sector_costs({"lenv": tl, "psi": tp}, steps, "lenv", 2, dtype)
        degs = [int(d) for d in tl.indexes[2].degs()]
        for world in (2, 5):
            want_cuts = sh.row_line_cuts(cost, degs, world, snap=8)
            assert ref.b200_cut_rows(lenv, psi, ([0], [0]), 2, world, 8) == want_cuts
            for r in range(world):
                want = sh.restrict_tensor(tl, 2, want_cuts[r])
                got = ref.b200_row_slab(lenv, 2, want_cuts[r], want.indexes)
                expect = ref.RefTensor.from_bst(want)
                assert got.indexes_equal(expect)                       # the reference's own Index comparison
                g = got.to_bst()
                assert g.same_structure(want) and np.array_equal(g.data, want.data)
        for axis in (0, 1):                                            # other axes, arbitrary ranges
            ranges = [(0, int(d)) if i % 2 else (int(d) // 3, int(d)) for i, d in enumerate(tl.indexes[axis].degs())]
            want = sh.restrict_tensor(tl, axis, ranges)
            g = ref.b200_row_slab(lenv, axis, ranges, want.indexes).to_bst()
            assert g.same_structure(want) and np.array_equal(g.data, want.data)


def test_restricted_operand_is_a_row_slice():
    ts = make_tensors(64, np.complex128, 2)
    mine, info = sh.shard_heff_tensors(ts, 4, 1)
    lenv, sub = ts["lenv"], mine["lenv"]
    dense_full = lenv.to_dense()
    starts = np.concatenate([[0], np.cumsum(lenv.indexes[2].degs())[:-1]])
    rows = np.concatenate([np.arange(int(starts[s]) + lo, int(starts[s]) + hi) for s, (lo, hi) in enumerate(info.sector_ranges[1]) if hi > lo]).astype(np.int64)
    assert np.array_equal(sub.to_dense(), dense_full[:, :, rows])


def _worker(rank, world, port, D, q):
    import torch.distributed as dist
    from oracle import contract_np as onp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    ts = make_tensors(D, np.float64, 3)
    mine, info = sh.shard_heff_tensors(ts, world, rank)
    cur = dict(mine)
    for lhs, rhs, axes, out in wl.HEFF_STEPS:      # local work unit (oracle stands in for the CUDA kernels)
        cur[out] = onp.contract_np(cur[lhs], cur[rhs], axes)
    stride = max(info.local_elems)
    local = torch.zeros(stride, dtype=torch.float64)
    local[:info.local_elems[rank]] = torch.from_numpy(cur["out"].data)
    gathered = torch.zeros(world * stride, dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, local)
    full = np.zeros(info.full_elems)
    sh.unpack_slabs(info, gathered.numpy(), stride, full)
    if rank == 0:
        ref = dict(ts)
        for lhs, rhs, axes, out in wl.HEFF_STEPS:
            ref[out] = onp.contract_np(ref[lhs], ref[rhs], axes)
        q.put(float(np.linalg.norm(full - ref["out"].data) / np.linalg.norm(ref["out"].data)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_allgather_rebuilds_full_result():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 24, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-12


def test_plan_partition_host_only():
    """qlb200_plan_partition on a host-only plan: row slabs of all ranks tile C (no GPU needed)."""
    ts = make_tensors(300, np.float64, 4)
    m = tk.Match(ts["lenv"], ts["psi"], ([0], [0]))
    world = 5
    cov = np.zeros(m.c_elems, np.int32)
    for r in range(world):
        p = tk.ContractionPlan(None, m, np.float64)
        p.partition(world, r)
        off, ln = p.c_ranges()
        for o, l in zip(off, ln):
            cov[int(o):int(o + l)] += 1
        p.close()
    assert np.all(cov == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_sharded_chain_on_one_gpu_matches_unsharded(ctx, dtype):
    """Each rank's work unit run in turn on one device; slabs assembled == unsharded apply."""
    from tensortoolkit_b200.heff import ContractionChain
    ts = make_tensors(200, dtype, 5)
    full_chain = ContractionChain(ctx, ts, wl.HEFF_STEPS, dtype)
    full_chain.apply_device()
    want = full_chain.result("out").data
    full_chain.close()
    world = 3
    got = np.zeros_like(want)
    for r in range(world):
        mine, info = sh.shard_heff_tensors(ts, world, r)
        ch = ContractionChain(ctx, mine, wl.HEFF_STEPS, dtype)
        ch.apply_device()
        loc = ch.result("out").data
        assert loc.size == info.local_elems[r]
        for s in info.slabs[r]:
            got[s.full_offset:s.full_offset + s.length] = loc[s.local_offset:s.local_offset + s.length]
        ch.close()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12


@pytest.mark.gpu
def test_sharded_chain_world1_unpack(ctx):
    """ShardedChain with world=1 exercises the packed->full batched copy kernel."""
    import torch
    from tensortoolkit_b200.heff import ContractionChain, ShardedChain
    ts = make_tensors(150, np.complex128, 6)
    full_chain = ContractionChain(ctx, ts, wl.HEFF_STEPS, np.complex128)
    full_chain.apply_device()
    want = full_chain.result("out").data
    full_chain.close()
    st = torch.cuda.Stream()
    ctx.sync()
    sc = ShardedChain(ctx, ts, wl.HEFF_STEPS, "lenv", 2, np.complex128, 1, 0, exchange="allgather")
    sc.apply()
    ctx.sync(); torch.cuda.synchronize()
    got = sc.full.cpu().numpy()
    sc.close()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_fused_exchange_writes_every_replica(ctx, dtype):
    """The fused exchange on one device: every rank's last-step GEMM stores its row slabs into THREE
    replicas of the full result (its own + two stand-ins for NVLink peers) from inside the kernel
    epilogue; after all ranks ran, every replica equals the unsharded apply."""
    from tensortoolkit_b200.heff import ContractionChain, ShardedChain, DeviceBuffer
    ts = make_tensors(220, dtype, 9)
    full_chain = ContractionChain(ctx, ts, wl.HEFF_STEPS, dtype)
    full_chain.apply_device()
    want = full_chain.result("out").data
    full_chain.close()
    world = 3
    nbytes = want.nbytes
    stride = (nbytes + 255) & ~255
    replicas = [DeviceBuffer(ctx, 2 * stride) for _ in range(2)]      # double-buffered result: two halves per replica
    chains = []
    for r in range(world):
        sc = ShardedChain(ctx, ts, wl.HEFF_STEPS, "lenv", 2, dtype, world, r, exchange="fused", peers=[b.ptr for b in replicas])
        chains.append(sc)
    # rank 0's own buffer is the third replica: make ranks 1, 2 write into it as well
    for sc in chains[1:]:
        sc.peer_ptrs.append(chains[0].full_base)
    for half in (0, 1, 0):                 # consecutive applies alternate between the two halves of every replica
        for b in replicas + [chains[0].full_buf]:
            zero = np.zeros(stride // 8, np.float64)
            tk._lib.check(tk._lib.lib.qlb200_memcpy_h2d(ctx.h, b.ptr + half * stride, zero.ctypes.data, zero.nbytes), "h2d")
        ctx.sync()
        for sc in chains:
            assert sc.parity == half
            sc.apply()
        ctx.sync()
        assert chains[0].full_ptr == chains[0].full_base + half * stride
        for ptr in [chains[0].full_ptr] + [b.ptr + half * stride for b in replicas]:
            got = np.empty_like(want)
            tk._lib.check(tk._lib.lib.qlb200_memcpy_d2h(ctx.h, got.ctypes.data, ptr, got.nbytes), "d2h")
            ctx.sync()
            assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12
    for sc in chains:
        sc.close()
    for b in replicas:
        b.free()


@pytest.mark.gpu
def test_sharded_chain_with_idle_ranks(ctx):
    """D=40 over 8 ranks leaves some ranks without rows: they build no plans, launch nothing, and the others still
    rebuild the full result (all ranks store into one buffer here)."""
    from tensortoolkit_b200.heff import ContractionChain, ShardedChain
    ts = make_tensors(40, np.float64, 10)
    full_chain = ContractionChain(ctx, ts, wl.HEFF_STEPS, np.float64)
    full_chain.apply_device()
    want = full_chain.result("out").data
    full_chain.close()
    world = 8
    chains = [ShardedChain(ctx, ts, wl.HEFF_STEPS, "lenv", 2, np.float64, world, r, exchange="fused", peers=[]) for r in range(world)]
    assert any(sc.idle for sc in chains) and not all(sc.idle for sc in chains)
    for sc in chains[1:]:
        sc.peer_ptrs = [chains[0].full_base]
    for sc in chains:
        n = sc.apply()
        assert (n == 0) == sc.idle
    ctx.sync()
    got = np.empty_like(want)
    tk._lib.check(tk._lib.lib.qlb200_memcpy_d2h(ctx.h, got.ctypes.data, chains[0].full_ptr, got.nbytes), "d2h")
    ctx.sync()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12
    for sc in chains:
        sc.close()


def test_plan_reads_heff_blocks_in_place():
    """Host-only plans of the U(1) H_eff chain: with one-dimensional physical sectors every block's
    permutation is trivial or one 2-D transposition (after re-ordering the contracted axes), so the
    complex plan sends nothing through the permute kernel; PERMUTE_ALL restores the reference's
    block-by-block transposes (global_operations.h:922-964)."""
    from tensortoolkit_b200 import _lib
    ts = make_tensors(96, np.complex128, 4)
    shells = dict(ts)
    moved_default, moved_all = 0, 0
    for lhs, rhs, axes, out in wl.HEFF_STEPS:
        m = tk.Match(shells[lhs], shells[rhs], axes)
        shells[out] = m.result_shell(np.complex128)
        for flags, acc in ((_lib.PLAN_DETERMINISTIC, "d"), (_lib.PLAN_DETERMINISTIC | _lib.PLAN_PERMUTE_ALL, "a")):
            p = tk.ContractionPlan(None, m, np.complex128, flags)
            s = p.stats()
            if acc == "d":
                moved_default += s.permute_elems_a + s.permute_elems_b
            else:
                moved_all += s.permute_elems_a + s.permute_elems_b
            p.close()
    assert moved_default == 0
    assert moved_all > 0


@pytest.mark.gpu
def test_comm_ranks_through_the_c_abi_only():
    """tests/cpp/comm_ranks: N threads, one GPU each, NO torch / NCCL / Python in the data path -- qlb200_comm_* (symmetric
    buffers over CUDA VMM, fd passing over unix sockets, multicast mapping, device barrier), qlb200_fanout_copy,
    qlb200_plan_partition + qlb200_execute_bcast / _mcast.  One rank where the box has one GPU, all of them otherwise."""
    import subprocess
    import torch
    exe = os.path.join(os.path.dirname(__file__), "cpp", "comm_ranks")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/comm_ranks not built (make -C tensortoolkit_b200/csrc comm_test)")
    n = min(torch.cuda.device_count(), 8)
    for world in sorted({1, min(n, 2), n}):
        out = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr
