// tests/cpp/comm_ranks.cc -- the multi-GPU path through the C ABI alone (no torch, no NCCL, no Python): N threads, one GPU
// and one qlb200 context each, bootstrap all-gather = a mutex + condition variable.  Every rank
//   1. creates a communicator (qlb200_comm_create) and symmetric buffers (qlb200_comm_alloc: unicast peer pointers and,
//      where the fabric has it, the NVSwitch multicast mapping),
//   2. uploads 1/N of an input vector and fans it out to every GPU (qlb200_fanout_copy), barrier (qlb200_comm_barrier),
//   3. plans a ragged grouped contraction (qlb200_plan_create_raw), keeps its cost-balanced share of the output rows
//      (qlb200_plan_partition) and stores its tiles into the result buffer of EVERY GPU from the GEMM epilogue
//      (qlb200_execute_bcast, then qlb200_execute_mcast), barrier,
//   4. compares its replica of the full result with the unpartitioned contraction it ran alone.
// usage: comm_ranks [N=2]   -- exit code 0 and "PASS" on success.  Built by `make -C tensortoolkit_b200/csrc comm_test`.
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "qlb200.h"

namespace {

struct Gather {            // all-gather between threads: everybody deposits, the last one releases the round
  std::mutex mu;
  std::condition_variable cv;
  int world = 1, arrived = 0;
  unsigned long long round = 0;
  std::vector<char> buf, out;
};
struct RankArg { Gather *g; int rank; };

int AllGather(void *user, const void *send, void *recv, size_t bytes) {
  RankArg *a = static_cast<RankArg *>(user);
  Gather &g = *a->g;
  std::unique_lock<std::mutex> lk(g.mu);
  if (g.buf.size() < bytes * g.world) g.buf.resize(bytes * g.world);
  std::memcpy(g.buf.data() + bytes * a->rank, send, bytes);
  const unsigned long long my_round = g.round;
  if (++g.arrived == g.world) {
    g.out.assign(g.buf.begin(), g.buf.begin() + bytes * g.world);
    g.arrived = 0; ++g.round;
    g.cv.notify_all();
  } else {
    g.cv.wait(lk, [&] { return g.round != my_round; });
  }
  std::memcpy(recv, g.out.data(), bytes * g.world);
  return 0;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    int rc_ = (call);                                                                              \
    if (rc_ != QLB200_OK) { std::fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, #call, rc_, qlb200_last_error()); return 1; } \
  } while (0)

struct Table {
  std::vector<uint32_t> a_shape, b_shape;
  std::vector<uint64_t> a_off, b_off;
  std::vector<qlb200_task> tasks;
  uint64_t a_elems = 0, b_elems = 0, c_elems = 0;
};

Table MakeTable(unsigned seed) {      // ragged blocks, several pairs per output block, sizes around the tile edges
  std::mt19937_64 rng(seed);
  auto pick = [&](int lo, int hi) { return int(lo + rng() % uint64_t(hi - lo + 1)); };
  Table t;
  for (int c = 0; c < 48; ++c) {
    const int m = pick(5, 300), n = pick(3, 260), pairs = pick(1, 3);
    for (int p = 0; p < pairs; ++p) {
      const int k = pick(4, 500);
      qlb200_task tk;
      std::memset(&tk, 0, sizeof(tk));
      tk.a_ord = tk.b_ord = uint32_t(t.tasks.size()); tk.c_ord = uint32_t(c);
      tk.a_off = t.a_elems; tk.b_off = t.b_elems; tk.c_off = t.c_elems;
      tk.m = uint32_t(m); tk.k = uint32_t(k); tk.n = uint32_t(n); tk.sign = (c + p) % 4 == 0 ? -1 : 1; tk.first = p == 0;
      t.a_shape.push_back(m); t.a_shape.push_back(k); t.b_shape.push_back(k); t.b_shape.push_back(n);
      t.a_off.push_back(t.a_elems); t.b_off.push_back(t.b_elems);
      t.a_elems += uint64_t(m) * k; t.b_elems += uint64_t(k) * n;
      t.tasks.push_back(tk);
    }
    t.c_elems += uint64_t(m) * n;
  }
  return t;
}

double RelErr(const std::vector<double> &x, const std::vector<double> &y) {
  double d = 0, n = 0;
  for (size_t i = 0; i < x.size(); ++i) { d += (x[i] - y[i]) * (x[i] - y[i]); n += y[i] * y[i]; }
  return std::sqrt(d / (n > 0 ? n : 1));
}

int RunRank(Gather *g, int world, int rank, double *worst) {
  RankArg arg{g, rank};
  qlb200_ctx *ctx = nullptr;
  CK(qlb200_ctx_create(rank, &ctx));
  qlb200_comm *comm = nullptr;
  CK(qlb200_comm_create(ctx, world, rank, AllGather, &arg, &comm));
  const bool has_mc = qlb200_comm_has_multicast(comm) != 0;

  const Table tb = MakeTable(20260017);
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<double> A(tb.a_elems), B(tb.b_elems);
  for (double &v : A) v = U(rng);
  for (double &v : B) v = U(rng);
  void *dA, *dB, *dC;
  CK(qlb200_dev_alloc(ctx, tb.a_elems * 8, &dA));
  CK(qlb200_dev_alloc(ctx, tb.b_elems * 8, &dB));
  CK(qlb200_dev_alloc(ctx, tb.c_elems * 8, &dC));
  CK(qlb200_memcpy_h2d(ctx, dB, B.data(), tb.b_elems * 8));

  // ---- input fan-out: rank r uploads its 1/world share of A and stores it into every GPU's copy ----
  void *symA = nullptr, *mcA = nullptr;
  std::vector<void *> peersA(world);
  const size_t a_bytes = (tb.a_elems * 8 + 255) & ~size_t(255);
  CK(qlb200_comm_alloc(comm, a_bytes, &symA, peersA.data(), &mcA));
  const uint64_t per = ((tb.a_elems + world - 1) / world + 1) & ~1ull;          // 16-byte granules
  const uint64_t lo = std::min<uint64_t>(tb.a_elems, per * rank), hi = std::min<uint64_t>(tb.a_elems, per * (rank + 1));
  if (hi > lo) {
    CK(qlb200_memcpy_h2d(ctx, static_cast<char *>(symA) + lo * 8, A.data() + lo, (hi - lo) * 8));
    const uint64_t nbytes = ((hi - lo) * 8 + 15) & ~15ull;
    if (has_mc && mcA != nullptr) CK(qlb200_fanout_copy(ctx, symA, lo * 8, nbytes, nullptr, 0, mcA));
    else CK(qlb200_fanout_copy(ctx, symA, lo * 8, nbytes, peersA.data(), world, nullptr));
  }
  CK(qlb200_comm_barrier(comm));
  std::vector<double> Aback(tb.a_elems);
  CK(qlb200_memcpy_d2h(ctx, Aback.data(), symA, tb.a_elems * 8));
  CK(qlb200_ctx_sync(ctx));
  if (std::memcmp(Aback.data(), A.data(), tb.a_elems * 8) != 0) { std::fprintf(stderr, "rank %d: fanned-out input differs\n", rank); return 1; }
  dA = symA;      // the contraction reads the fanned-out copy

  // ---- the whole contraction alone (the answer), then this rank's share stored into every GPU's result ----
  const int32_t ident[2] = {0, 1};
  qlb200_plan *whole = nullptr, *mine = nullptr;
  CK(qlb200_plan_create_raw(ctx, QLB200_F64, QLB200_PLAN_DETERMINISTIC, 2, ident, tb.a_off.size(), tb.a_shape.data(), tb.a_off.data(), 2, ident,
                            tb.b_off.size(), tb.b_shape.data(), tb.b_off.data(), tb.tasks.size(), tb.tasks.data(), tb.c_elems, &whole));
  CK(qlb200_execute(ctx, whole, dA, dB, dC, QLB200_MEM_DEVICE));
  std::vector<double> want(tb.c_elems), got(tb.c_elems);
  CK(qlb200_memcpy_d2h(ctx, want.data(), dC, tb.c_elems * 8));
  CK(qlb200_ctx_sync(ctx));
  CK(qlb200_plan_create_raw(ctx, QLB200_F64, QLB200_PLAN_DETERMINISTIC | QLB200_PLAN_STAGGER_OUTPUT, 2, ident, tb.a_off.size(), tb.a_shape.data(),
                            tb.a_off.data(), 2, ident, tb.b_off.size(), tb.b_shape.data(), tb.b_off.data(), tb.tasks.size(), tb.tasks.data(),
                            tb.c_elems, &mine));
  CK(qlb200_plan_partition(mine, world, rank));
  void *symC = nullptr, *mcC = nullptr;
  std::vector<void *> peersC(world);
  CK(qlb200_comm_alloc(comm, (tb.c_elems * 8 + 255) & ~size_t(255), &symC, peersC.data(), &mcC));
  for (int pass = 0; pass < (has_mc && mcC != nullptr ? 2 : 1); ++pass) {
    // own pointer first (qlb200_execute_bcast's convention), then the peers'
    std::vector<void *> order;
    order.push_back(peersC[rank]);
    for (int p = 0; p < world; ++p) if (p != rank) order.push_back(peersC[p]);
    if (pass == 0) CK(qlb200_execute_bcast(ctx, mine, dA, dB, order.data(), world));
    else CK(qlb200_execute_mcast(ctx, mine, dA, dB, mcC));
    CK(qlb200_comm_barrier(comm));
    CK(qlb200_memcpy_d2h(ctx, got.data(), symC, tb.c_elems * 8));
    CK(qlb200_ctx_sync(ctx));
    const double err = RelErr(got, want);
    if (err > *worst) *worst = err;
    if (!(err <= 1e-12)) { std::fprintf(stderr, "rank %d: %s result differs, rel err %.3e\n", rank, pass ? "multicast" : "peer-store", err); return 1; }
    // clear before the next pass; the barrier keeps a fast rank's stores out of a slow rank's clear
    std::vector<double> zero(tb.c_elems, 0.0);
    CK(qlb200_comm_barrier(comm));
    CK(qlb200_memcpy_h2d(ctx, symC, zero.data(), tb.c_elems * 8));
    CK(qlb200_ctx_sync(ctx));
    CK(qlb200_comm_barrier(comm));
    CK(qlb200_ctx_sync(ctx));
  }
  qlb200_plan_destroy(whole);
  qlb200_plan_destroy(mine);
  qlb200_dev_free(ctx, dB); qlb200_dev_free(ctx, dC);
  qlb200_comm_destroy(comm);
  qlb200_ctx_destroy(ctx);
  if (rank == 0) std::printf("multicast mapping: %s\n", has_mc ? "yes" : "no");
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  if (world < 1 || world > 8) { std::fprintf(stderr, "usage: comm_ranks [1..8]\n"); return 2; }
  Gather g;
  g.world = world;
  std::vector<int> rc(world, 0);
  std::vector<double> worst(world, 0.0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back([&, r] { rc[r] = RunRank(&g, world, r, &worst[r]); });
  for (auto &t : th) t.join();
  double w = 0;
  for (int r = 0; r < world; ++r) { if (rc[r] != 0) { std::printf("FAIL (rank %d)\n", r); return 1; } w = std::max(w, worst[r]); }
  std::printf("PASS: %d ranks, sharded == whole to %.2e on every replica (C ABI only: qlb200_comm_* + qlb200_execute_bcast / _mcast)\n", world, w);
  return 0;
}
