"""oracle/contract_np.py -- TEST INFRASTRUCTURE ONLY (the "port" oracle).

A numpy restatement of the reference's CPU algorithm for qlten::Contract, written to be read next
to the reference, not to be fast.  Pinned against the reference itself: tests/test_oracle.py checks
it against oracle/_ref/libqlref.so (the unmodified reference compiled here) and against the golden
vectors under tests/golden/ that were dumped from that library (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Each function cites the reference code it restates (paths relative to /root/reference/include/qlten).
"""
import numpy as np

from tensortoolkit_b200.tensor import BlockSparseTensor


def saved_axes(rank_a, rank_b, axes):
    """TenCtrctGenSavedAxesSet -- qltensor/blk_spar_data_ten/data_blk_operations.h:268-296."""
    return ([i for i in range(rank_a) if i not in axes[0]], [i for i in range(rank_b) if i not in axes[1]])


def trans_orders(axes, saved):
    """TenCtrctNeedTransCheck -- tensor_manipulation/ten_ctrct.h:396-443.
    A: saved ++ contracted, B: contracted ++ saved; a transpose is needed iff not the identity."""
    pa = list(saved[0]) + list(axes[0])
    pb = list(axes[1]) + list(saved[1])
    return pa, pa != sorted(pa), pb, pb != sorted(pb)


def fermion_exchange_sign(a_par, b_par, a_ctrct, b_ctrct, a_ctrct_dirs):
    """FermionExchangeSignForCtrct -- data_blk_operations.h:351-401 (restated with the same
    rotate-and-renumber bookkeeping)."""
    par = list(a_par) + list(b_par)
    ra = len(a_par)
    cat = []
    for x, y in zip(a_ctrct, b_ctrct):
        cat += [x, y + ra]
    exch = 0
    for i in range(len(a_ctrct)):
        ax1, ax2 = cat[2 * i], cat[2 * i + 1]
        p1, p2 = par[ax1], par[ax2]
        assert p1 == p2
        if p2:
            exch += sum(par[ax1 + 1:ax2])
        # std::rotate(begin+ax1+1, begin+ax2, begin+ax2+1): element at ax2 moves to ax1+1
        par = par[:ax1 + 1] + [par[ax2]] + par[ax1 + 1:ax2] + par[ax2 + 1:]
        cat = [c + 1 if ax1 < c < ax2 else c for c in cat]
        cat[2 * i + 1] = ax1 + 1
        exch += int(bool(p1) and a_ctrct_dirs[i] == -1)
    return -1 if exch & 1 else 1


def fermionic_reorder_sign(parities, order):
    """FermionicInplaceReorder -- utility/utils_inl.h:53-75 (cycle-following swaps)."""
    par = list(parities)
    ind = list(order)
    exch = 0
    for i in range(len(ind)):
        cur = i
        while i != ind[cur]:
            nxt = ind[cur]
            between = sum(par[min(cur, nxt) + 1:max(cur, nxt)])
            exch += between * (par[cur] + par[nxt]) + par[cur] * par[nxt]
            par[cur], par[nxt] = par[nxt], par[cur]
            ind[cur] = cur
            cur = nxt
        ind[cur] = cur
    return -1 if exch & 1 else 1


def _blk_parities(t: BlockSparseTensor, b: int):
    k = t.kind
    return [k.parity(t.indexes[i].sectors[int(c)].qn) for i, c in enumerate(t.blk_coors[b])]


def match_tasks(a: BlockSparseTensor, b: BlockSparseTensor, axes, saved=None):
    """DataBlkGenForTenCtrct -- data_blk_operations.h:411-578: scan every (a, b) block pair in
    ascending blk_idx order, keep those whose coordinates agree on the contracted axes.
    `saved` overrides the order of the free axes in the result (the contiguous-axes executor passes a
    cyclic order, contract_contiguous_axes.h:346-352).
    Returns (tasks in discovery order, c_coors sorted by blk_idx, c_offsets)."""
    sa, sb = saved if saved is not None else saved_axes(a.rank, b.rank, axes)
    c_nsct = [a.indexes[i].nsct for i in sa] + [b.indexes[i].nsct for i in sb]
    scalar = len(c_nsct) == 0
    fermi = a.kind.fermionic
    dirs = [a.indexes[x].dir for x in axes[0]]
    tasks, c_seen = [], {}
    for i in range(a.nblk):
        ka = tuple(int(a.blk_coors[i, x]) for x in axes[0])
        for j in range(b.nblk):
            if ka != tuple(int(b.blk_coors[j, y]) for y in axes[1]):
                continue
            m = int(np.prod([int(a.blk_shape[i, x]) for x in sa], dtype=np.int64)) if sa else 1
            k = int(np.prod([int(a.blk_shape[i, x]) for x in axes[0]], dtype=np.int64)) if len(axes[0]) else 1
            n = int(np.prod([int(b.blk_shape[j, y]) for y in sb], dtype=np.int64)) if sb else 1
            cc = tuple(int(a.blk_coors[i, x]) for x in sa) + tuple(int(b.blk_coors[j, y]) for y in sb)
            cidx = 0
            for c, ns in zip(cc, c_nsct):
                cidx = cidx * ns + c
            if scalar:
                beta = 0.0 if not tasks else 1.0       # :486-489, first task overwrites
            else:
                beta = 1.0 if cidx in c_seen else 0.0  # :523-543
                if cidx not in c_seen:
                    shape = tuple(int(a.blk_shape[i, x]) for x in sa) + tuple(int(b.blk_shape[j, y]) for y in sb)
                    c_seen[cidx] = (cc, shape)
            sign = 1
            if fermi:
                sign = fermion_exchange_sign(_blk_parities(a, i), _blk_parities(b, j), axes[0], axes[1], dirs)
            tasks.append(dict(a=i, b=j, a_idx=int(a.blk_idx[i]), b_idx=int(b.blk_idx[j]), c_idx=cidx,
                              a_off=int(a.blk_offset[i]), b_off=int(b.blk_offset[j]), m=m, k=k, n=n, sign=sign, beta=beta))
            if scalar:
                break                                  # :478 one B block per A block
    c_keys = sorted(c_seen)
    c_off, off = {}, 0
    for key in c_keys:                                 # DataBlksOffsetRefresh, :137-145
        c_off[key] = off
        off += int(np.prod(c_seen[key][1], dtype=np.int64))
    for t in tasks:
        t["c_off"] = 0 if scalar else c_off[t["c_idx"]]
    c_coors = np.array([c_seen[k][0] for k in c_keys], dtype=np.uint32).reshape(len(c_keys), len(c_nsct))
    return tasks, c_coors, (1 if (scalar and tasks) else off)


def contract_np(a: BlockSparseTensor, b: BlockSparseTensor, axes) -> BlockSparseTensor:
    """Contract -- ten_ctrct.h:277-290 -> CtrctTwoBSDTAndAssignIn, global_operations.h:895-992:
    per task, transpose each operand block (TensorTranspose, framework/hp_numeric/ten_trans.h:130-186:
    output axis j = input axis perm[j]) and C = sign * A' B' + beta * C (blas_level3.h:35-108)."""
    axes = (list(axes[0]), list(axes[1]))
    sa, sb = saved_axes(a.rank, b.rank, axes)
    pa, _, pb, _ = trans_orders(axes, (sa, sb))
    tasks, c_coors, c_elems = match_tasks(a, b, axes)
    dt = np.result_type(a.dtype, b.dtype)
    c = BlockSparseTensor([a.indexes[i] for i in sa] + [b.indexes[i] for i in sb], dt)
    if c.rank:
        c.set_blocks(c_coors)
    c.data = np.zeros(c_elems, dtype=dt)
    # SortTasksByCBlkIdx (raw_data_operation_tasks.h:250-269): by C block, beta = 0 first
    for t in sorted(tasks, key=lambda t: (t["c_idx"], t["beta"])):
        am = np.transpose(a.block(t["a"]), pa).reshape(t["m"], t["k"])
        bm = np.transpose(b.block(t["b"]), pb).reshape(t["k"], t["n"])
        out = c.data[t["c_off"]:t["c_off"] + t["m"] * t["n"]].reshape(t["m"], t["n"])
        prod = t["sign"] * (am.astype(dt) @ bm.astype(dt))
        if t["beta"] == 0.0:
            out[...] = prod
        else:
            out += prod
    return c


def residue_fermion_sign(parities, saved, trans_critical_axe):
    """CountResidueFermionSignForMatBasedCtrct -- data_blk_operations.h:579-611, for one block."""
    first = [ax for ax in saved if ax < trans_critical_axe]
    second = [ax for ax in saved if ax >= trans_critical_axe]
    if not first or not second:
        return 1
    n1 = sum(parities[:len(first)])
    n2 = sum(parities[second[0]:])
    return -1 if (n1 & 1) and (n2 & 1) else 1


def contract_contiguous_np(a: BlockSparseTensor, b: BlockSparseTensor, a_start: int, b_start: int, size: int) -> BlockSparseTensor:
    """ContractContiguousAxes<Tail, Head> -- tensor_manipulation/contract_contiguous_axes.h:849-873 ->
    MatrixBasedTensorContractionExecutor: GenerateDataBlk_ (:333-364: contracted axes (start + i) % rank, free axes in
    cyclic order from the end of the contracted range), TransposePrepare_ (:473-534: the blocks are rotated as
    matrices about the critical axis -- Tail: end of A's range, Head: start of B's -- and fermionic blocks pick up the
    residue sign), then C = sign * A' B' per task (CtrctAccordingTask)."""
    ra, rb = a.rank, b.rank
    axes = ([(a_start + i) % ra for i in range(size)], [(b_start + i) % rb for i in range(size)])
    a_end, b_end = (a_start + size) % ra, (b_start + size) % rb
    sa = [(a_end + i) % ra for i in range(ra - size)]
    sb = [(b_end + i) % rb for i in range(rb - size)]
    a_crit, b_crit = a_end, b_start                     # <Tail, Head>, :325-326
    tasks, c_coors, c_elems = match_tasks(a, b, axes, saved=(sa, sb))
    dt = np.result_type(a.dtype, b.dtype)
    c = BlockSparseTensor([a.indexes[i] for i in sa] + [b.indexes[i] for i in sb], dt)
    if c.rank:
        c.set_blocks(c_coors)
    c.data = np.zeros(c_elems, dtype=dt)
    fermi = a.kind.fermionic
    # matrix rotation about the critical axis: axes [crit, rank) come first, then [0, crit)
    rot_a = list(range(a_crit, ra)) + list(range(a_crit))
    rot_b = list(range(b_crit, rb)) + list(range(b_crit))
    for t in sorted(tasks, key=lambda t: (t["c_idx"], t["beta"])):
        am = np.transpose(a.block(t["a"]), rot_a).reshape(t["m"], t["k"])
        bm = np.transpose(b.block(t["b"]), rot_b).reshape(t["k"], t["n"])
        sign = t["sign"]
        if fermi:
            if a_crit > 0:
                sign *= residue_fermion_sign(_blk_parities(a, t["a"]), sa, a_crit)
            if b_crit > 0:
                sign *= residue_fermion_sign(_blk_parities(b, t["b"]), sb, b_crit)
        out = c.data[t["c_off"]:t["c_off"] + t["m"] * t["n"]].reshape(t["m"], t["n"])
        prod = sign * (am.astype(dt) @ bm.astype(dt))
        if t["beta"] == 0.0:
            out[...] = prod
        else:
            out += prod
    return c


class AccumulateLayoutMismatchNp(ValueError):
    pass


def contract_accumulate_np(a: BlockSparseTensor, b: BlockSparseTensor, a_start: int, b_start: int, size: int, alpha, beta,
                           c: BlockSparseTensor = None, allow_expand: bool = True) -> BlockSparseTensor:
    """ContractTailHeadContiguousAccumulate -- tensor_manipulation/contract_contiguous_axes.h:954-1000 (executor
    :333-475 GenerateDataBlk_, :567-640 ExpandOutputTopology_, :642-680 ScaleUntouchedOutputBlocks_, :682-782):
        c <- beta * c + alpha * ContractContiguousAxes(a, b)
    on the UNION of c's blocks and the blocks the contraction produces; c None = default tensor (beta must be 0);
    blocks missing from c need allow_expand (the Try... probe forbids it)."""
    prod = contract_contiguous_np(a, b, a_start, b_start, size)
    if c is None:
        if beta != 0:
            raise ValueError("beta must be 0 for a default output")
        out = BlockSparseTensor(prod.indexes, prod.dtype)
        if prod.rank == 0:
            out.data = alpha * prod.data if prod.data.size else np.zeros(1, prod.dtype)
            return out
        if prod.nblk:
            out.set_blocks(prod.blk_coors, alpha * prod.data)
        return out
    if list(c.indexes) != list(prod.indexes):
        raise AccumulateLayoutMismatchNp("indexes")
    out = BlockSparseTensor(prod.indexes, prod.dtype)
    if prod.rank == 0:
        old = c.data[0] if c.data.size else 0.0
        new = prod.data[0] if prod.data.size else 0.0
        out.data = np.array([beta * old + alpha * new], dtype=prod.dtype)
        return out
    old_keys = {tuple(int(x) for x in c.blk_coors[i]): i for i in range(c.nblk)}
    new_keys = {tuple(int(x) for x in prod.blk_coors[i]): i for i in range(prod.nblk)}
    if not allow_expand and any(k not in old_keys for k in new_keys):
        raise AccumulateLayoutMismatchNp("output block topology requires expansion")
    keys = sorted(set(old_keys) | set(new_keys))
    if not keys:
        return out
    out.set_blocks(np.array(keys, np.uint32).reshape(len(keys), prod.rank))
    out.data[...] = 0
    for nb in range(out.nblk):
        k = tuple(int(x) for x in out.blk_coors[nb])
        blk = out.block(nb)
        if k in old_keys and beta != 0:
            blk += beta * c.block(old_keys[k])
        if k in new_keys:
            blk += alpha * prod.block(new_keys[k])
    return out


def apply_rank2_axes_np(x: BlockSparseTensor, ops) -> BlockSparseTensor:
    """dmrg::ApplyRank2ToAxisPreserveOrder / ApplyTwoRank2ToAxesPreserveOrder -- tensor_manipulation/dmrg/axis_ops.h:2889-3125.
    `ops` = [(rank-2 op, axis), ...] (one or two, distinct axes); op in {input_index, output_index} layout.  For every input
    block and every op block whose input sector matches the block's sector on that axis (Rank2OpBlockIndicesByInputSector
    :960-982), the output block with the axis coordinates replaced (:2950-2953, :3078-3081) accumulates
    in[.., i1, .., i2, ..] * op1[i1, j1] * op2[i2, j2] (AddRank2AxisBlock :1695-1775, AddTwoRank2AxesBlockGemm :1932-1986)."""
    idxs = list(x.indexes)
    for op, ax in ops:
        idxs[ax] = op.indexes[1]
    out = BlockSparseTensor(idxs, x.dtype)
    by_sector = []
    for op, ax in ops:
        d = {}
        for b in range(op.nblk):
            d.setdefault(int(op.blk_coors[b, 0]), []).append(b)
        by_sector.append(d)
    contrib = {}
    for ib in range(x.nblk):
        lists = [by_sector[o].get(int(x.blk_coors[ib, ax]), []) for o, (op, ax) in enumerate(ops)]
        combos = [(b1,) for b1 in lists[0]] if len(ops) == 1 else [(b1, b2) for b1 in lists[0] for b2 in lists[1]]
        for combo in combos:
            c = [int(v) for v in x.blk_coors[ib]]
            for (op, ax), b in zip(ops, combo):
                c[ax] = int(op.blk_coors[b, 1])
            contrib.setdefault(tuple(c), []).append((ib, combo))
    if not contrib:
        return out
    out.set_blocks(np.array(sorted(contrib), np.uint32).reshape(len(contrib), x.rank))
    out.data[...] = 0
    for nb in range(out.nblk):
        acc = out.block(nb)
        for ib, combo in contrib[tuple(int(v) for v in out.blk_coors[nb])]:
            v = x.block(ib)
            for (op, ax), b in zip(ops, combo):
                v = np.moveaxis(np.tensordot(v, op.block(b), axes=([ax], [0])), -1, ax)
            acc += v
    return out


def transpose_np(t: BlockSparseTensor, order) -> BlockSparseTensor:
    """QLTensor::Transpose -- qltensor/qltensor_impl.h:449-464 -> BlockSparseDataTensor::Transpose,
    global_operations.h:393-441: every block permuted, fermionic blocks scaled by the reorder sign
    (DataBlk::Transpose, data_blk.h:124-146), new blk_idx/offsets from the permuted coordinates."""
    order = list(order)
    out = BlockSparseTensor([t.indexes[i] for i in order], t.dtype)
    if t.nblk == 0:
        return out
    new_coors = t.blk_coors[:, order]
    out.set_blocks(new_coors)
    # map new (sorted) blocks back to the source blocks
    key = {tuple(int(x) for x in new_coors[b]): b for b in range(t.nblk)}
    for nb in range(out.nblk):
        src = key[tuple(int(x) for x in out.blk_coors[nb])]
        sign = fermionic_reorder_sign(_blk_parities(t, src), order) if t.kind.fermionic else 1
        out.block(nb)[...] = sign * np.transpose(t.block(src), order)
    return out


def estimate_cost(a: BlockSparseTensor, b: BlockSparseTensor, axes) -> dict:
    """EstimateContractCost -- tensor_manipulation/tensor_op_cost.h:467-516."""
    axes = (list(axes[0]), list(axes[1]))
    sa, sb = saved_axes(a.rank, b.rank, axes)
    _, ta, _, tb = trans_orders(axes, (sa, sb))
    tasks, c_coors, c_elems = match_tasks(a, b, axes)
    s = a.dtype.itemsize
    f = 8.0 if a.dtype == np.complex128 else 2.0
    cost = dict(flops=0.0, gemm_count=len(tasks), candidate_block_pair_count=a.nblk * b.nblk,
                output_block_count=len(c_coors), output_raw_elem_count=(c_elems if len(c_coors) else 0),
                read_bytes=0, write_bytes=0, temp_peak_bytes=0)
    for t in tasks:
        cost["flops"] += f * t["m"] * t["k"] * t["n"]
        cost["read_bytes"] += (t["m"] * t["k"] + t["k"] * t["n"]) * s + (t["m"] * t["n"] * s if t["beta"] != 0.0 else 0)
        cost["write_bytes"] += t["m"] * t["n"] * s
    temp = 0
    if ta:
        temp += sum(int(a.blk_size[i]) for i in {t["a"] for t in tasks}) * s
    if tb:
        temp += sum(int(b.blk_size[j]) for j in {t["b"] for t in tasks}) * s
    cost["temp_peak_bytes"] = temp
    cost["read_bytes"] += temp
    cost["write_bytes"] += temp
    return cost
