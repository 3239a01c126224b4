// matcher.cc -- quantum-number sector matcher: pairs compatible blocks of two block-sparse tensors
// on the host and emits the packed task table + output block structure.
//
// Behavioural contract (must reproduce bit-exactly, checked by tests/test_matcher.py against the
// reference's own task list): BlockSparseDataTensor::DataBlkGenForTenCtrct,
// reference include/qlten/qltensor/blk_spar_data_ten/data_blk_operations.h:411-578, and its
// filtered variant DataBlkGenFor1SectTenCtrct (data_blk_operator_separate_ctrct.h:29-160).
//
// Design differences (B200-first host side, not a port):
//  * the reference scans all N_A x N_B block pairs comparing 64-bit hashes of the contracted
//    coordinates (never verifying them); here B blocks are bucketed by the EXACT mixed-radix key of
//    their contracted coordinates, so matching is O(N_A + N_B log N_B + pairs) and collision-free (dense key / block-index tables when the sector spaces are small).  Buckets
//    keep ascending block order, so the discovery order (and therefore which pair is "first" for a
//    C block) is the reference's.
//  * everything is flat arrays; no per-hit std::map lookups or vector allocations.
#include "matcher.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace qlb200 {

std::string Shell::Load(const qlb200_shell *s) {
  if (s == nullptr) return "null shell";
  if (s->rank < 0 || s->rank > QLB200_MAX_RANK) return "rank out of range";
  rank = s->rank;
  if (rank > 0 && (s->nsct == nullptr || s->deg == nullptr)) return "shell without sector tables";
  if (s->nblk > 0 && rank > 0 && s->blk_coors == nullptr) return "shell with blocks but no block coordinates";
  nsct.assign(s->nsct, s->nsct + rank);
  sct_base.resize(rank + 1);
  uint32_t tot = 0;
  for (int i = 0; i < rank; ++i) {
    if (nsct[i] == 0) return "index without sectors";
    sct_base[i] = tot;
    tot += nsct[i];
  }
  sct_base[rank] = tot;
  deg.assign(s->deg, s->deg + tot);
  if (s->parity != nullptr) parity.assign(s->parity, s->parity + tot); else parity.clear();
  if (s->dir != nullptr) dir.assign(s->dir, s->dir + rank); else dir.assign(rank, 0);
  nblk = s->nblk;
  coors.assign(s->blk_coors, s->blk_coors + nblk * rank);
  shape.resize(nblk * rank);
  size.resize(nblk);
  offset.resize(nblk);
  blk_idx.resize(nblk);
  elems = 0;
  uint64_t prev_idx = 0;
  for (uint64_t b = 0; b < nblk; ++b) {
    uint64_t sz = 1, idx = 0;
    for (int i = 0; i < rank; ++i) {
      uint32_t c = coors[b * rank + i];
      if (c >= nsct[i]) return "block coordinate out of range";
      uint32_t d = deg[sct_base[i] + c];
      shape[b * rank + i] = d;
      sz *= d;
      idx = idx * nsct[i] + c;
    }
    if (b > 0 && idx <= prev_idx) return "blocks must be listed in strictly ascending blk_idx order";
    prev_idx = idx;
    size[b] = sz;
    offset[b] = elems;
    blk_idx[b] = idx;
    elems += sz;
  }
  return "";
}

int FermionCtrctSign(const uint8_t *a_par, int a_rank, const uint8_t *b_par, int b_rank,
                     const std::vector<int> &a_ctrct, const std::vector<int> &b_ctrct,
                     const int8_t *a_dir) {
  // Legs of A followed by legs of B form one ordered list of (possibly odd) objects.  Each
  // contracted pair is brought together by sliding B's leg leftwards until it sits right behind
  // its partner on A; every odd leg it jumps over contributes a transposition when the pair
  // itself is odd.  An odd pair whose A leg is IN (ket) contributes one more.
  int legs[2 * QLB200_MAX_RANK];
  uint8_t par[2 * QLB200_MAX_RANK];
  const int total = a_rank + b_rank;
  for (int i = 0; i < a_rank; ++i) { legs[i] = i; par[i] = a_par[i]; }
  for (int i = 0; i < b_rank; ++i) { legs[a_rank + i] = a_rank + i; par[a_rank + i] = b_par[i]; }
  int swaps = 0;
  for (size_t p = 0; p < a_ctrct.size(); ++p) {
    const int la = a_ctrct[p], lb = a_rank + b_ctrct[p];
    int pa = -1, pb = -1;
    for (int i = 0; i < total; ++i) {
      if (legs[i] == la) pa = i;
      if (legs[i] == lb) pb = i;
    }
    const bool odd = par[pb] != 0;
    if (odd) {
      for (int i = pa + 1; i < pb; ++i) swaps += par[i];
      if (a_dir[la] == QLB200_DIR_IN) swaps += 1;
    }
    // slide lb from pb to pa + 1
    const int leg = legs[pb];
    const uint8_t lp = par[pb];
    for (int i = pb; i > pa + 1; --i) { legs[i] = legs[i - 1]; par[i] = par[i - 1]; }
    legs[pa + 1] = leg;
    par[pa + 1] = lp;
  }
  return (swaps & 1) ? -1 : 1;
}

int FermionReorderSign(const uint8_t *par, int rank, const int32_t *perm) {
  // new leg j is old leg perm[j]; the sign is the parity of the permutation restricted to the
  // odd-parity legs = number of inverted odd pairs.
  int inv = 0;
  for (int i = 0; i < rank; ++i) {
    if (!par[perm[i]]) continue;
    for (int j = i + 1; j < rank; ++j) {
      if (par[perm[j]] && perm[i] > perm[j]) ++inv;
    }
  }
  return (inv & 1) ? -1 : 1;
}

std::vector<qlb200_task> Match::SortedTasks() const {
  std::vector<qlb200_task> t = tasks;
  std::stable_sort(t.begin(), t.end(), [](const qlb200_task &x, const qlb200_task &y) {
    if (x.c_blk_idx != y.c_blk_idx) return x.c_blk_idx < y.c_blk_idx;
    return x.first > y.first;
  });
  return t;
}

namespace {

// Sign the contiguous-axes executor applies per operand block (reference: data_blk_operations.h:579-611,
// CountResidueFermionSignForMatBasedCtrct): the free legs of the block fall into the group in front of the
// contracted range and the group behind it; bringing the rear group to the front (the cyclic order of the
// result) costs a sign when both groups hold an odd number of odd-parity legs.
int ResidueSign(const uint8_t *par, const std::vector<int> &saved, const std::vector<int> &ctrct) {
  if (ctrct.empty()) return 1;
  const int first_ctrct = ctrct.front();
  int n_front = 0, n_rear = 0;
  for (int ax : saved) (ax < first_ctrct ? n_front : n_rear) += par[ax] ? 1 : 0;
  return ((n_front & 1) && (n_rear & 1)) ? -1 : 1;
}

}  // namespace

std::string BuildMatch(const qlb200_shell *sa, const qlb200_shell *sb, int nctrct, const int32_t *a_axes,
                       const int32_t *b_axes, int sel_axis, uint32_t sel_sector, Match *out,
                       const int32_t *a_saved_order, const int32_t *b_saved_order, bool mat_based_residue) {
  Match &m = *out;
  std::string err = m.a.Load(sa);
  if (!err.empty()) return "A: " + err;
  err = m.b.Load(sb);
  if (!err.empty()) return "B: " + err;
  const Shell &A = m.a, &B = m.b;
  if (A.rank == 0 || B.rank == 0) return "scalar operands are not contractible";
  if (nctrct < 0 || nctrct > A.rank || nctrct > B.rank) return "bad number of contracted axes";
  if (A.fermionic() != B.fermionic()) return "A and B disagree on fermionic-ness";
  std::vector<char> a_used(A.rank, 0), b_used(B.rank, 0);
  for (int i = 0; i < nctrct; ++i) {
    int x = a_axes[i], y = b_axes[i];
    if (x < 0 || x >= A.rank || y < 0 || y >= B.rank) return "contracted axis out of range";
    if (a_used[x] || b_used[y]) return "axis contracted twice";
    a_used[x] = b_used[y] = 1;
    if (A.nsct[x] != B.nsct[y]) return "contracted indexes have different sector counts";
    for (uint32_t s = 0; s < A.nsct[x]; ++s) {
      if (A.deg[A.sct_base[x] + s] != B.deg[B.sct_base[y] + s]) return "contracted indexes have different degeneracies";
    }
    m.a_ctrct.push_back(x);
    m.b_ctrct.push_back(y);
  }
  if (sel_axis >= 0) {
    if (sel_axis >= A.rank || a_used[sel_axis]) return "1-sector axis must be a free axis of A";
    if (sel_sector >= A.nsct[sel_axis]) return "1-sector index out of range";
  }
  auto fill_saved = [](int rank, const std::vector<char> &used, const int32_t *order, std::vector<int> *saved) -> std::string {
    if (order == nullptr) {
      for (int i = 0; i < rank; ++i) if (!used[i]) saved->push_back(i);
      return "";
    }
    std::vector<char> seen(rank, 0);
    int nfree = 0;
    for (int i = 0; i < rank; ++i) nfree += used[i] ? 0 : 1;
    for (int i = 0; i < nfree; ++i) {
      const int ax = order[i];
      if (ax < 0 || ax >= rank || used[ax] || seen[ax]) return "saved-axes order must list every free axis exactly once";
      seen[ax] = 1;
      saved->push_back(ax);
    }
    return "";
  };
  err = fill_saved(A.rank, a_used, a_saved_order, &m.a_saved);
  if (!err.empty()) return "A: " + err;
  err = fill_saved(B.rank, b_used, b_saved_order, &m.b_saved);
  if (!err.empty()) return "B: " + err;
  m.a_perm = m.a_saved; m.a_perm.insert(m.a_perm.end(), m.a_ctrct.begin(), m.a_ctrct.end());
  m.b_perm = m.b_ctrct; m.b_perm.insert(m.b_perm.end(), m.b_saved.begin(), m.b_saved.end());
  m.a_need_trans = !std::is_sorted(m.a_perm.begin(), m.a_perm.end());
  m.b_need_trans = !std::is_sorted(m.b_perm.begin(), m.b_perm.end());
  m.c_rank = static_cast<int>(m.a_saved.size() + m.b_saved.size());
  if (m.c_rank > QLB200_MAX_RANK) return "result rank exceeds QLB200_MAX_RANK";
  m.scalar = (m.c_rank == 0);
  for (int ax : m.a_saved) m.c_nsct.push_back(A.nsct[ax]);
  for (int ax : m.b_saved) m.c_nsct.push_back(B.nsct[ax]);
  m.candidate_pairs = A.nblk * B.nblk;

  // exact mixed-radix key of the contracted coordinates
  {
    unsigned __int128 span = 1;
    for (int ax : m.a_ctrct) span *= A.nsct[ax];
    if (span > (static_cast<unsigned __int128>(1) << 63)) return "contracted sector space too large";
  }
  auto key_of = [](const Shell &S, uint64_t blk, const std::vector<int> &axes, const Shell &radix_shell,
                   const std::vector<int> &radix_axes) {
    uint64_t k = 0;
    for (size_t i = 0; i < axes.size(); ++i) k = k * radix_shell.nsct[radix_axes[i]] + S.coors[blk * S.rank + axes[i]];
    return k;
  };
  // bucket B blocks by key, ascending block order inside a bucket (counting sort over sorted keys)
  std::vector<std::pair<uint64_t, uint32_t>> b_keys(B.nblk);
  for (uint64_t j = 0; j < B.nblk; ++j) b_keys[j] = {key_of(B, j, m.b_ctrct, A, m.a_ctrct), static_cast<uint32_t>(j)};
  std::sort(b_keys.begin(), b_keys.end());
  // key -> [begin, end) in b_keys: a dense table over the contracted sector space when that is small (the usual
  // case: a few bond sectors per contracted index), else binary search in the sorted key list
  uint64_t key_span = 1;
  for (int ax : m.a_ctrct) key_span *= A.nsct[ax];
  constexpr uint64_t kDenseLimit = 1u << 22;
  // QLB200_MATCH_NO_DENSE=1 forces the large-sector-space code paths (tests)
  const bool allow_dense = std::getenv("QLB200_MATCH_NO_DENSE") == nullptr;
  const bool dense_keys = allow_dense && key_span <= kDenseLimit;
  std::vector<uint32_t> bucket_begin;      // [key_span + 1] when dense
  if (dense_keys) {
    bucket_begin.assign(key_span + 1, 0);
    for (const auto &kv : b_keys) ++bucket_begin[kv.first + 1];
    for (uint64_t k = 0; k < key_span; ++k) bucket_begin[k + 1] += bucket_begin[k];
  }
  auto find_bucket = [&](uint64_t key, uint32_t *lo, uint32_t *hi) {
    if (dense_keys) { *lo = bucket_begin[key]; *hi = bucket_begin[key + 1]; return *hi > *lo; }
    auto first = std::lower_bound(b_keys.begin(), b_keys.end(), std::make_pair(key, uint32_t(0)));
    auto last = first;
    while (last != b_keys.end() && last->first == key) ++last;
    *lo = uint32_t(first - b_keys.begin()); *hi = uint32_t(last - b_keys.begin());
    return *hi > *lo;
  };

  // c_blk_idx -> "already created": a bitmap over C's block-index space when that is small, else a hash set
  unsigned __int128 c_span128 = 1;
  for (uint32_t n : m.c_nsct) c_span128 *= n;
  const bool dense_c = allow_dense && c_span128 <= (static_cast<unsigned __int128>(1) << 27);
  std::vector<uint64_t> c_bitmap(dense_c ? (static_cast<uint64_t>(c_span128) + 63) / 64 : 0, 0);
  std::unordered_map<uint64_t, uint32_t> c_seen;
  std::vector<CBlock> c_unsorted;
  uint8_t a_par[QLB200_MAX_RANK], b_par[QLB200_MAX_RANK];
  const bool fermi = A.fermionic();
  // fermion exchange sign depends only on the two blocks' leg-parity patterns: memoise per (mask_a, mask_b)
  std::vector<int8_t> sign_cache;
  if (fermi) sign_cache.assign(size_t(1) << (A.rank + B.rank), 0);
  m.tasks.reserve(A.nblk + B.nblk);
  // per-B-block quantities that do not depend on the partner: n, the B part of the C block index, the parity mask
  const size_t n_a_saved = m.a_saved.size();
  uint64_t b_radix = 1;                         // product of the sector counts of B's free axes
  for (int ax : m.b_saved) b_radix *= B.nsct[ax];
  std::vector<uint64_t> b_nn(B.nblk), b_cpart(B.nblk);
  std::vector<uint32_t> b_masks(fermi ? B.nblk : 0);
  for (uint64_t j = 0; j < B.nblk; ++j) {
    const uint32_t *bc = &B.coors[j * B.rank];
    const uint32_t *bsh = &B.shape[j * B.rank];
    uint64_t nn = 1, part = 0;
    for (int ax : m.b_saved) { nn *= bsh[ax]; part = part * B.nsct[ax] + bc[ax]; }
    b_nn[j] = nn; b_cpart[j] = part;
    if (fermi) {
      uint32_t mask = 0;
      for (int r = 0; r < B.rank; ++r) mask |= uint32_t(B.parity[B.sct_base[r] + bc[r]] != 0) << r;
      b_masks[j] = mask;
    }
  }
  for (uint64_t i = 0; i < A.nblk; ++i) {
    const uint32_t *ac = &A.coors[i * A.rank];
    if (sel_axis >= 0 && ac[sel_axis] != sel_sector) continue;
    uint32_t q_lo, q_hi;
    if (!find_bucket(key_of(A, i, m.a_ctrct, A, m.a_ctrct), &q_lo, &q_hi)) continue;
    const uint32_t *ash = &A.shape[i * A.rank];
    uint64_t mm = 1, kk = 1, a_cpart = 0;
    for (int ax : m.a_saved) { mm *= ash[ax]; a_cpart = a_cpart * A.nsct[ax] + ac[ax]; }
    for (int ax : m.a_ctrct) kk *= ash[ax];
    if (mm > UINT32_MAX || kk > UINT32_MAX) return "block dimension exceeds 2^32";
    a_cpart *= b_radix;
    uint32_t a_mask = 0;
    if (fermi) for (int r = 0; r < A.rank; ++r) { a_par[r] = A.parity[A.sct_base[r] + ac[r]]; a_mask |= uint32_t(a_par[r] != 0) << r; }
    for (uint32_t q = q_lo; q < q_hi; ++q) {
      const uint64_t j = b_keys[q].second;
      const uint64_t nn = b_nn[j];
      if (nn > UINT32_MAX) return "block dimension exceeds 2^32";
      m.tasks.emplace_back();
      qlb200_task &t = m.tasks.back();
      std::memset(&t, 0, sizeof(t));
      t.a_blk_idx = A.blk_idx[i]; t.b_blk_idx = B.blk_idx[j];
      t.a_off = A.offset[i]; t.b_off = B.offset[j];
      t.a_ord = static_cast<uint32_t>(i); t.b_ord = static_cast<uint32_t>(j);
      t.k = static_cast<uint32_t>(kk);
      t.sign = 1;
      if (m.scalar) {
        // all axes contracted: one B block can match; C is the size-1 raw buffer
        t.m = t.n = 1;
        t.c_blk_idx = 0; t.c_ord = 0;
        t.first = m.tasks.size() == 1 ? 1 : 0;
      } else {
        t.m = static_cast<uint32_t>(mm); t.n = static_cast<uint32_t>(nn);
        const uint64_t cidx = a_cpart + b_cpart[j];
        t.c_blk_idx = cidx;
        bool fresh;
        if (dense_c) {
          uint64_t &w = c_bitmap[cidx >> 6];
          const uint64_t bit = uint64_t(1) << (cidx & 63);
          fresh = !(w & bit);
          w |= bit;
        } else {
          fresh = c_seen.emplace(cidx, static_cast<uint32_t>(c_unsorted.size())).second;
        }
        t.first = fresh ? 1 : 0;
        if (fresh) {
          const uint32_t *bc = &B.coors[j * B.rank];
          const uint32_t *bsh = &B.shape[j * B.rank];
          c_unsorted.emplace_back();
          CBlock &cb = c_unsorted.back();
          std::memset(&cb, 0, sizeof(cb));
          size_t r = 0;
          for (int ax : m.a_saved) { cb.coors[r] = ac[ax]; cb.shape[r] = ash[ax]; ++r; }
          for (int ax : m.b_saved) { cb.coors[r] = bc[ax]; cb.shape[r] = bsh[ax]; ++r; }
          (void) n_a_saved;
          cb.blk_idx = cidx; cb.size = mm * nn;
        }
      }
      if (fermi) {
        int8_t &sg = sign_cache[(size_t(b_masks[j]) << A.rank) | a_mask];
        if (sg == 0) {
          const uint32_t *bc = &B.coors[j * B.rank];
          for (int r = 0; r < B.rank; ++r) b_par[r] = B.parity[B.sct_base[r] + bc[r]];
          int v = FermionCtrctSign(a_par, A.rank, b_par, B.rank, m.a_ctrct, m.b_ctrct, A.dir.data());
          if (mat_based_residue) v *= ResidueSign(a_par, m.a_saved, m.a_ctrct) * ResidueSign(b_par, m.b_saved, m.b_ctrct);
          sg = static_cast<int8_t>(v);
        }
        t.sign = sg;
      }
      if (m.scalar) break;
    }
  }

  if (m.scalar) {
    m.c_elems = m.tasks.empty() ? 0 : 1;
    return "";
  }
  // offsets by prefix sum in ascending blk_idx order (DataBlksOffsetRefresh)
  std::sort(c_unsorted.begin(), c_unsorted.end(), [](const CBlock &x, const CBlock &y) { return x.blk_idx < y.blk_idx; });
  uint64_t off = 0;
  for (size_t i = 0; i < c_unsorted.size(); ++i) {
    c_unsorted[i].offset = off;
    off += c_unsorted[i].size;
  }
  m.c_blocks = std::move(c_unsorted);
  m.c_elems = off;
  // ordinal of a C block = its rank among the created blocks: popcount prefix over the bitmap, or binary search
  std::vector<uint32_t> word_rank;
  if (dense_c) {
    word_rank.resize(c_bitmap.size() + 1, 0);
    for (size_t w = 0; w < c_bitmap.size(); ++w) word_rank[w + 1] = word_rank[w] + uint32_t(__builtin_popcountll(c_bitmap[w]));
  }
  for (auto &t : m.tasks) {
    if (dense_c) {
      const uint64_t w = t.c_blk_idx >> 6, bit = t.c_blk_idx & 63;
      t.c_ord = word_rank[w] + uint32_t(__builtin_popcountll(c_bitmap[w] & ((uint64_t(1) << bit) - 1)));
    } else {
      auto it = std::lower_bound(m.c_blocks.begin(), m.c_blocks.end(), t.c_blk_idx,
                                 [](const CBlock &cb, uint64_t v) { return cb.blk_idx < v; });
      t.c_ord = uint32_t(it - m.c_blocks.begin());
    }
    t.c_off = m.c_blocks[t.c_ord].offset;
  }
  return "";
}

void EstimateCost(const Match &m, int dtype, qlb200_cost *out) {
  std::memset(out, 0, sizeof(*out));
  const uint64_t s = (dtype == QLB200_C64) ? 16 : 8;
  const double fl = (dtype == QLB200_C64) ? 8.0 : 2.0;
  out->candidate_block_pair_count = m.candidate_pairs;
  out->output_block_count = m.scalar ? 0 : m.c_blocks.size();
  out->output_raw_elem_count = m.scalar ? 0 : m.c_elems;
  std::vector<char> a_seen(m.a.nblk, 0), b_seen(m.b.nblk, 0);
  uint64_t temp = 0;
  for (const auto &t : m.tasks) {
    const uint64_t mm = t.m, kk = t.k, nn = t.n;
    out->flops += fl * static_cast<double>(mm) * static_cast<double>(kk) * static_cast<double>(nn);
    out->gemm_count += 1;
    out->read_bytes += (mm * kk + kk * nn) * s;
    if (!t.first) out->read_bytes += mm * nn * s;
    out->write_bytes += mm * nn * s;
    if (m.a_need_trans && !a_seen[t.a_ord]) { a_seen[t.a_ord] = 1; temp += m.a.size[t.a_ord] * s; }
    if (m.b_need_trans && !b_seen[t.b_ord]) { b_seen[t.b_ord] = 1; temp += m.b.size[t.b_ord] * s; }
  }
  out->temp_peak_bytes = temp;
  out->read_bytes += temp;
  out->write_bytes += temp;
}

}  // namespace qlb200
