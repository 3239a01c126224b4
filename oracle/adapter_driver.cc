// oracle/adapter_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// The drop-in adapter (include/qlten_b200/contract.h) instantiated on the reference's own QLTensor types, behind a C
// ABI: this is how the parity tests call the CUDA path exactly like a TensorToolkit user would
// (qlten::b200::Contract(&A, &B, axes, &C)).  Handles are the TenBase* objects of oracle/ref_handles.h, created by
// libqlref.so.  Built by oracle/Makefile into oracle/_ref/libqladapter.so; links tensortoolkit_b200/libqlb200.so.
#include "ref_handles.h"

#include "qlten_b200/contract.h"
#include "qlten_b200/axis_ops.h"
#include "qlten_b200/sharding.h"

using namespace qlref;

namespace {

// Calls f(const TenBox<ElemT, QNT>*) for the element / quantum-number type the handle was created with; every
// instantiation of the generic lambda returns the same type R.
template<typename R, typename F>
R Dispatch(const TenBase *t, F &&f) {
#define QLREF_CASE(K, QN)                                                                   \
  case K:                                                                                   \
    if (t->dtype == D_F64) return f(static_cast<const TenBox<QLTEN_Double, QN> *>(t));      \
    return f(static_cast<const TenBox<QLTEN_Complex, QN> *>(t));
  switch (t->kind) {
    QLREF_CASE(K_U1, U1QN)
    QLREF_CASE(K_FU1, fU1QN)
    QLREF_CASE(K_U1U1, U1U1QN)
    QLREF_CASE(K_FU1U1, fU1U1QN)
    QLREF_CASE(K_Z2, Z2QN)
    QLREF_CASE(K_FZ2, fZ2QN)
    default: break;
  }
#undef QLREF_CASE
  throw std::invalid_argument("unknown quantum-number kind in handle");
}

template<typename Box> const Box *Same(const Box *, const TenBase *o) { return static_cast<const Box *>(o); }

template<typename Fn>
void *Guard(const char *what, Fn &&fn) {
  try { return fn(); }
  catch (const std::exception &e) { std::fprintf(stderr, "%s: %s\n", what, e.what()); return nullptr; }
}

}  // namespace

extern "C" {

void *qlref_b200_contract(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba, void *ctx) {
  return Guard("qlref_b200_contract", [&]() -> void * {
    return Dispatch<TenBase *>(static_cast<const TenBase *>(a), [&](auto *A) -> TenBase * {
      typename std::remove_pointer_t<decltype(A)>::Ten c;
      qlten::b200::Contract(&A->t, &Same(A, static_cast<const TenBase *>(b))->t, MakeAxes(n, aa, ba), &c, (qlb200_ctx *) ctx);
      return A->wrap(std::move(c));
    });
  });
}

void *qlref_b200_contract_1sector(const void *a, int64_t axis, int64_t sct, const void *b, int n, const int64_t *aa,
                                  const int64_t *ba, void *ctx) {
  return Guard("qlref_b200_contract_1sector", [&]() -> void * {
    return Dispatch<TenBase *>(static_cast<const TenBase *>(a), [&](auto *A) -> TenBase * {
      typename std::remove_pointer_t<decltype(A)>::Ten c;
      qlten::b200::Contract1Sector(&A->t, (size_t) axis, (size_t) sct, &Same(A, static_cast<const TenBase *>(b))->t,
                                   MakeAxes(n, aa, ba), &c, (qlb200_ctx *) ctx);
      return A->wrap(std::move(c));
    });
  });
}

// side: 0 = <Tail, Head> (default), 1 = <Head, Head>, 2 = <Tail, Tail>, 3 = <Head, Tail>
void *qlref_b200_contract_contiguous(const void *a, const void *b, int64_t a_start, int64_t b_start, int64_t size, int side,
                                     void *ctx) {
  return Guard("qlref_b200_contract_contiguous", [&]() -> void * {
    return Dispatch<TenBase *>(static_cast<const TenBase *>(a), [&](auto *A) -> TenBase * {
      using Box = std::remove_pointer_t<decltype(A)>;
      using E = typename Box::Elem;
      using Q = typename Box::QN;
      typename Box::Ten c;
      const auto &tb = Same(A, static_cast<const TenBase *>(b))->t;
      qlb200_ctx *cx = (qlb200_ctx *) ctx;
      switch (side) {
        case 1: qlten::b200::ContractContiguousAxes<E, Q, CtrctSide::Head, CtrctSide::Head>(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c, cx); break;
        case 2: qlten::b200::ContractContiguousAxes<E, Q, CtrctSide::Tail, CtrctSide::Tail>(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c, cx); break;
        case 3: qlten::b200::ContractContiguousAxes<E, Q, CtrctSide::Head, CtrctSide::Tail>(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c, cx); break;
        default: qlten::b200::ContractContiguousAxes<E, Q, CtrctSide::Tail, CtrctSide::Head>(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size, c, cx); break;
      }
      return A->wrap(std::move(c));
    });
  });
}

// returns 1 = done, 0 = the Try... probe reported a layout mismatch, -1 = exception (text on stderr)
int qlref_b200_contract_accumulate(const void *a, const void *b, int64_t a_start, int64_t b_start, int64_t size, const double *alpha2,
                                   const double *beta2, void *c, int try_only, uint64_t *stats10, void *ctx) {
  try {
    return Dispatch<int>(static_cast<const TenBase *>(a), [&](auto *A) -> int {
      using Box = std::remove_const_t<std::remove_pointer_t<decltype(A)>>;
      auto &tc = static_cast<Box *>(static_cast<TenBase *>(c))->t;
      const auto &tb = Same(A, static_cast<const TenBase *>(b))->t;
      ContiguousContractStats st;
      int ok = 1;
      if (try_only) {
        ok = qlten::b200::TryContractTailHeadContiguousAccumulate(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size,
                                                                  Box::MakeScalar(alpha2), Box::MakeScalar(beta2), tc, &st, (qlb200_ctx *) ctx) ? 1 : 0;
      } else {
        qlten::b200::ContractTailHeadContiguousAccumulate(A->t, tb, (size_t) a_start, (size_t) b_start, (size_t) size,
                                                          Box::MakeScalar(alpha2), Box::MakeScalar(beta2), tc, &st, (qlb200_ctx *) ctx);
      }
      if (stats10) Box::StoreStats(st, stats10);
      return ok;
    });
  } catch (const std::exception &e) { std::fprintf(stderr, "qlref_b200_contract_accumulate: %s\n", e.what()); return -1; }
}

void *qlref_b200_apply_rank2(const void *x, const void *op1, int64_t axis1, const void *op2, int64_t axis2, void *ctx) {
  return Guard("qlref_b200_apply_rank2", [&]() -> void * {
    return Dispatch<TenBase *>(static_cast<const TenBase *>(x), [&](auto *X) -> TenBase * {
      using Box = std::remove_const_t<std::remove_pointer_t<decltype(X)>>;
      if constexpr (Fermionicable<typename Box::QN>::IsFermionic()) {
        return nullptr;
      } else {
        typename Box::Ten out;
        const auto &o1 = Same(X, static_cast<const TenBase *>(op1))->t;
        if (op2 == nullptr) qlten::b200::dmrg::ApplyRank2ToAxisPreserveOrder(X->t, o1, (size_t) axis1, out, (qlb200_ctx *) ctx);
        else qlten::b200::dmrg::ApplyTwoRank2ToAxesPreserveOrder(X->t, o1, (size_t) axis1, Same(X, static_cast<const TenBase *>(op2))->t,
                                                                 (size_t) axis2, out, (qlb200_ctx *) ctx);
        return X->wrap(std::move(out));
      }
    });
  });
}

int qlref_b200_transpose(void *t, const int64_t *perm, void *ctx) {
  try {
    return Dispatch<int>(static_cast<const TenBase *>(t), [&](auto *A) -> int {
      using Box = std::remove_pointer_t<decltype(A)>;
      auto &ten = const_cast<typename std::remove_const_t<Box> &>(*A).t;
      std::vector<size_t> o(ten.Rank());
      for (size_t i = 0; i < ten.Rank(); ++i) o[i] = (size_t) perm[i];
      qlten::b200::Transpose(&ten, o, (qlb200_ctx *) ctx);
      return 0;
    });
  } catch (const std::exception &e) { std::fprintf(stderr, "qlref_b200_transpose: %s\n", e.what()); return -1; }
}

// qlten::b200::RowSlab: rows [lo_hi[2 s], lo_hi[2 s + 1]) of every sector s of index `axis` (host only)
void *qlref_b200_row_slab(const void *t, int64_t axis, int64_t nsct, const int64_t *lo_hi) {
  return Guard("qlref_b200_row_slab", [&]() -> void * {
    return Dispatch<TenBase *>(static_cast<const TenBase *>(t), [&](auto *T) -> TenBase * {
      qlten::b200::RowRanges ranges;
      for (int64_t s = 0; s < nsct; ++s) ranges.push_back({uint32_t(lo_hi[2 * s]), uint32_t(lo_hi[2 * s + 1])});
      return T->wrap(qlten::b200::RowSlab(T->t, size_t(axis), ranges));
    });
  });
}

// qlten::b200::SectorFlops + CutRowLine for one contraction: ranges_out[(r * nsct + s) * 2 + {0, 1}]
int qlref_b200_cut_rows(const void *a, const void *b, int n, const int64_t *aa, const int64_t *ba, int64_t split_axis, int world,
                        int snap, int64_t *ranges_out) {
  void *ok = Guard("qlref_b200_cut_rows", [&]() -> void * {
    return Dispatch<void *>(static_cast<const TenBase *>(a), [&](auto *A) -> void * {
      std::vector<double> cost;
      qlten::b200::SectorFlops(A->t, Same(A, static_cast<const TenBase *>(b))->t, MakeAxes(n, aa, ba), size_t(split_axis), cost);
      const auto &idx = A->t.GetIndex(size_t(split_axis));
      std::vector<uint32_t> degs;
      for (size_t s = 0; s < idx.GetQNSctNum(); ++s) degs.push_back(uint32_t(idx.GetQNSct(s).GetDegeneracy()));
      const auto cuts = qlten::b200::CutRowLine(cost, degs, world, snap);
      for (int r = 0; r < world; ++r)
        for (size_t s = 0; s < degs.size(); ++s) {
          ranges_out[(size_t(r) * degs.size() + s) * 2] = cuts[r][s].first;
          ranges_out[(size_t(r) * degs.size() + s) * 2 + 1] = cuts[r][s].second;
        }
      return const_cast<void *>(a);
    });
  });
  return ok ? 0 : 1;
}

}  // extern "C"
