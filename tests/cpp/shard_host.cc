// tests/cpp/shard_host.cc -- the HOST side of the multi-GPU path from a C++ client of the C ABI (no GPU, no Python at run
// time): read two tensor shells (lenv, psi) from a file, match lenv x psi (step 1 of the H_eff chain), attribute its flops to
// the sectors of lenv's free bond (qlb200_shard_sector_flops), cut the row line for `world` ranks (qlb200_shard_cut_line),
// build every rank's row slab of lenv (qlb200_shard_restrict), match the slab against psi and print what a rank would compute.
// tests/test_sharding.py writes the file and compares the numbers with tensortoolkit_b200/sharding.py.
// usage: shard_host FILE WORLD   (built by the test: g++ -std=c++17 -Iinclude ... -lqlb200)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "qlb200.h"

struct Shell {
  int32_t rank = 0;
  std::vector<uint32_t> nsct, deg, coors;
  std::vector<int8_t> dir;
  uint64_t nblk = 0;
  qlb200_shell view() const {
    qlb200_shell s;
    s.rank = rank; s.nsct = nsct.data(); s.deg = deg.data(); s.parity = nullptr; s.dir = dir.data(); s.nblk = nblk; s.blk_coors = coors.data();
    return s;
  }
};

static bool Read(FILE *f, Shell *s) {
  unsigned long long nblk;
  if (std::fscanf(f, "%d %llu", &s->rank, &nblk) != 2) return false;
  s->nblk = nblk;
  s->nsct.resize(s->rank); s->dir.resize(s->rank);
  size_t total = 0;
  for (int i = 0; i < s->rank; ++i) { int d; if (std::fscanf(f, "%u %d", &s->nsct[i], &d) != 2) return false; s->dir[i] = int8_t(d); total += s->nsct[i]; }
  s->deg.resize(total);
  for (size_t i = 0; i < total; ++i) if (std::fscanf(f, "%u", &s->deg[i]) != 1) return false;
  s->coors.resize(s->nblk * s->rank);
  for (size_t i = 0; i < s->coors.size(); ++i) if (std::fscanf(f, "%u", &s->coors[i]) != 1) return false;
  return true;
}

#define CK(call) do { if ((call) != QLB200_OK) { std::fprintf(stderr, "%s failed: %s\n", #call, qlb200_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = std::fopen(argv[1], "r");
  if (!f) return 2;
  Shell lenv, psi;
  if (!Read(f, &lenv) || !Read(f, &psi)) return 2;
  std::fclose(f);
  const int world = std::atoi(argv[2]);
  const int32_t a_axes[1] = {0}, b_axes[1] = {0};        // lenv[vb OUT, wb OUT, vb IN] x psi[vb IN, ph, ph, vb OUT] over the first bond
  const int32_t split_axis = 2;                          // lenv's free ket-side bond stays free through the chain
  qlb200_shell sl = lenv.view(), sp = psi.view();
  qlb200_match *m = nullptr;
  CK(qlb200_match_create(&sl, &sp, 1, a_axes, b_axes, &m));
  const uint32_t nsct = lenv.nsct[split_axis];
  std::vector<double> cost(nsct, 0.0);
  CK(qlb200_shard_sector_flops(m, split_axis, QLB200_C64, cost.data()));
  std::printf("whole: tasks %llu c_elems %llu\n", (unsigned long long) qlb200_match_ntask(m), (unsigned long long) qlb200_match_c_elems(m));
  qlb200_match_destroy(m);
  size_t base = 0;                                       // start of the split index's sectors in the index-major deg array
  for (int i = 0; i < split_axis; ++i) base += lenv.nsct[i];
  const uint32_t *sdeg = lenv.deg.data() + base;
  std::vector<qlb200_piece> line(nsct);
  for (uint32_t s = 0; s < nsct; ++s) { line[s].sector = s; line[s].lo = 0; line[s].hi = sdeg[s]; line[s].pad_ = 0; line[s].weight = sdeg[s] ? cost[s] / sdeg[s] : 0.0; }
  std::vector<uint32_t> ranges(size_t(world) * nsct * 2);
  CK(qlb200_shard_cut_line(line.data(), nsct, sdeg, nsct, world, 8, ranges.data()));
  for (int r = 0; r < world; ++r) {
    const uint32_t *rg = ranges.data() + size_t(r) * nsct * 2;
    qlb200_slab_info info;
    CK(qlb200_shard_restrict(&sl, split_axis, rg, &info, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    std::vector<uint32_t> kept(info.nsct_kept + 1), ndeg(info.nsct_kept + 1), kblk(info.nblk_kept + 1), ncoor((info.nblk_kept + 1) * lenv.rank);
    std::vector<uint64_t> cs(info.ncopy + 1), cd(info.ncopy + 1), cl(info.ncopy + 1);
    CK(qlb200_shard_restrict(&sl, split_axis, rg, &info, kept.data(), ndeg.data(), kblk.data(), ncoor.data(), cs.data(), cd.data(), cl.data()));
    unsigned long long rows = 0, tasks = 0, c_elems = 0;
    for (uint32_t s = 0; s < nsct; ++s) rows += rg[2 * s + 1] - rg[2 * s];
    if (info.nblk_kept > 0) {
      Shell slab = lenv;                                 // the slab's shell: kept sectors of the split index, kept blocks
      slab.nsct[split_axis] = info.nsct_kept;
      slab.deg.assign(lenv.deg.begin(), lenv.deg.begin() + base);                     // the indexes in front of the split one
      slab.deg.insert(slab.deg.end(), ndeg.begin(), ndeg.begin() + info.nsct_kept);
      slab.deg.insert(slab.deg.end(), lenv.deg.begin() + base + nsct, lenv.deg.end());   // ... and behind it
      slab.nblk = info.nblk_kept;
      slab.coors.assign(ncoor.begin(), ncoor.begin() + info.nblk_kept * lenv.rank);
      qlb200_shell ss = slab.view();
      qlb200_match *ms = nullptr;
      CK(qlb200_match_create(&ss, &sp, 1, a_axes, b_axes, &ms));
      tasks = qlb200_match_ntask(ms); c_elems = qlb200_match_c_elems(ms);
      qlb200_match_destroy(ms);
    }
    std::printf("rank %d: rows %llu slab_elems %llu copies %llu tasks %llu c_elems %llu\n", r, rows, (unsigned long long) info.elems,
                (unsigned long long) info.ncopy, tasks, c_elems);
  }
  return 0;
}
