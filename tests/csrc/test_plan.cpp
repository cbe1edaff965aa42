// Host-side invariants of the launch plan (compiled with g++ by tests/test_plan_host.py; no GPU needed).
//   1. piece_grid + PieceIter: for every launch geometry the pieces of all CTAs cover every (row block, column tile)
//      exactly once, pieces of a row block are numbered 0..npi-1 in column order, and no row block has more pieces than
//      the plan sized accumulator copies for.
//   2. build_plan: workspace regions do not overlap and respect the 1024-byte alignment; shape limits return the
//      documented codes.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../focal_b200/csrc/plan.h"

using namespace fb;

static int fails = 0;
#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      ++fails;                                             \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);     \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
      if (fails > 20) std::exit(1);                        \
    }                                                      \
  } while (0)

static void check_geometry(int n_items, int T, int sms, bool streamk) {
  int np = 0;
  const int grid = piece_grid(n_items, T, sms, streamk, &np);
  CHECK(grid >= 1 && grid <= sms, "grid %d for items=%d T=%d sms=%d", grid, n_items, T, sms);
  CHECK(np >= 1 && np <= kMaxPieces, "np %d", np);
  std::vector<int> cover((size_t)n_items * T, 0), next_pk(n_items, 0), npi_of(n_items, 0), next_t(n_items, 0);
  for (int cta = 0; cta < grid; ++cta) {
    PieceIter it(cta, grid, n_items, T, streamk);
    for (int item, t0, t1, pk, npi; it.next(item, t0, t1, pk, npi);) {
      CHECK(item >= 0 && item < n_items && t0 >= 0 && t0 < t1 && t1 <= T, "piece (%d, %d, %d)", item, t0, t1);
      CHECK(npi >= 1 && npi <= np, "npi %d > planned %d (items=%d T=%d sms=%d)", npi, np, n_items, T, sms);
      CHECK(pk == next_pk[item], "piece index %d, expected %d (item %d)", pk, next_pk[item], item);
      CHECK(t0 == next_t[item], "pieces of item %d not contiguous: t0 %d, expected %d", item, t0, next_t[item]);
      CHECK((pk == 0) == (t0 == 0), "primary piece must start at tile 0");
      CHECK(npi_of[item] == 0 || npi_of[item] == npi, "pieces of item %d disagree on their count", item);
      npi_of[item] = npi;
      ++next_pk[item];
      next_t[item] = t1;
      for (int t = t0; t < t1; ++t) ++cover[(size_t)item * T + t];
    }
  }
  for (int i = 0; i < n_items; ++i) {
    CHECK(next_pk[i] == npi_of[i], "item %d: %d pieces seen, %d announced", i, next_pk[i], npi_of[i]);
    for (int t = 0; t < T; ++t)
      CHECK(cover[(size_t)i * T + t] == 1, "tile (%d, %d) covered %d times (items=%d T=%d sms=%d sk=%d)", i, t,
            cover[(size_t)i * T + t], n_items, T, sms, (int)streamk);
  }
}

static FocalCfg cfg(int B, int S, int M, int D, int s0, int s1, int need_grad = 1) {
  FocalCfg c{};
  c.B = B; c.S = S; c.M = M; c.D = D; c.temperature = 0.5f; c.margin = 1.f;
  c.w_shared = 1; c.w_private = 1; c.w_orth = 3; c.w_rank = 5;
  c.need_grad = need_grad; c.terms = FOCAL_TERM_ALL; c.seq_begin = s0; c.seq_end = s1;
  return c;
}

static void check_plan(const FocalCfg& c, int sms) {
  Plan p;
  const int rc = build_plan(c, p, sms);
  CHECK(rc == FOCAL_OK, "build_plan rc %d (B=%d S=%d M=%d D=%d)", rc, c.B, c.S, c.M, c.D);
  if (rc) return;
  struct Region { uint64_t off, bytes; const char* name; };
  std::vector<Region> r;
  const uint64_t rowsNce = (uint64_t)p.S * p.bpad;
  for (int o = 0; o < p.nOps; ++o) r.push_back({p.ops[o].off, (uint64_t)p.ops[o].kb * rowsNce * 128, "operand"});
  r.push_back({p.xt_off, (uint64_t)p.nT * p.kbFull * p.Bpad * 128, "xt"});
  r.push_back({p.sq_off, (uint64_t)p.nT * p.Bpad * 4, "sq"});
  r.push_back({p.mintra_off, (uint64_t)p.nT * p.Bpad * 4, "mintra"});
  const uint64_t rs = (uint64_t)p.nProb * p.S * 2 * p.bpad * 4;
  r.push_back({p.rpart_off, rs * p.nsplit_fwd, "rpart"});
  r.push_back({p.rsum_off, rs, "rsum"});
  r.push_back({p.rinv_off, rs, "rinv"});
  for (int k = 0; k < p.np_nce; ++k) r.push_back({p.dz_off + k * p.dz2_delta, p.dz_bytes, "dz"});
  for (int k = 0; k < p.np_tmp; ++k) {
    r.push_back({p.dx_off + k * p.dx2_delta, p.dx_bytes, "dx"});
    r.push_back({p.rho_off + k * p.rho2_delta, (uint64_t)p.nT * p.Bpad * 4, "rho"});
    r.push_back({p.cnt_off + k * p.cnt2_delta, (uint64_t)p.nT * p.bpad * 4, "cnt"});
  }
  r.push_back({p.flag_tmp_off, (uint64_t)p.nT * (p.Bpad / kTileM) * 4, "flag_tmp"});
  r.push_back({p.flag_nce_off, (uint64_t)p.nProb * p.S * 2 * (p.bpad / kTileM) * 4, "flag_nce"});
  r.push_back({p.part1_off, (uint64_t)p.nblk1 * 16, "part1"});
  r.push_back({p.part2_off, (uint64_t)p.nblk2 * 8, "part2"});
  r.push_back({p.part3_off, (uint64_t)p.nitems3 * 4, "part3"});
  r.push_back({p.lossd_off, 64, "lossd"});
  r.push_back({p.bar_off, 64, "bar"});
  r.push_back({p.lossx_off, (uint64_t)kMaxPeers * 64, "lossx"});
  for (size_t i = 0; i < r.size(); ++i) {
    CHECK(r[i].off % 1024 == 0, "%s not 1024-aligned", r[i].name);
    CHECK(r[i].off + r[i].bytes <= p.total_bytes, "%s beyond the workspace", r[i].name);
    for (size_t j = 0; j < i; ++j)
      CHECK(r[i].off + r[i].bytes <= r[j].off || r[j].off + r[j].bytes <= r[i].off, "%s overlaps %s (B=%d D=%d M=%d)",
            r[i].name, r[j].name, c.B, c.D, c.M);
  }
  CHECK(p.bpad % kTileM == 0 && p.Bpad % kTileM == 0 && p.bpad >= p.b && p.Bpad >= p.Bt && p.Bt >= p.B, "padding");
  CHECK(p.Sp >= p.S && (p.Sp & (p.Sp - 1)) == 0 && p.Sp < 2 * p.S + (p.S == 1) && p.Bt == p.b * p.Sp, "temporal row space");
  // the launch geometries the plan stores must be the ones PieceIter will see
  const int bnT = tile_bn(p.kbFull);
  const int wide = (p.wide && p.need_grad) ? 2 : 1;
  check_geometry(p.nT * tmp_row_tiles(p) * wide, (p.Bt + bnT - 1) / bnT, sms, p.sk_tmp != 0);
  int np = 0;
  CHECK(piece_grid(p.nT * tmp_row_tiles(p) * wide, (p.Bt + bnT - 1) / bnT, sms, p.sk_tmp != 0, &np) == p.grid_tmp, "grid_tmp");
  CHECK(np <= p.np_tmp, "np_tmp %d < %d", p.np_tmp, np);
}

int main() {
  for (int sms : {1, 7, 132, 148})
    for (int T : {1, 2, 3, 16, 32, 86, 87, 103, 256})
      for (int n : {1, 2, 3, 5, 31, 32, 64, 100, 147, 148, 149, 256, 1024})
        for (int sk = 0; sk < 2; ++sk) check_geometry(n, T, sms, sk != 0);
  // headline, shards of it (2 / 4 / 8 ranks), the other BASELINE configs, edge shapes
  check_plan(cfg(8192, 4, 2, 256, 0, 2048), 148);
  for (int R : {2, 4, 8})
    for (int r = 0; r < R; ++r) check_plan(cfg(8192, 4, 2, 256, r * 2048 / R, (r + 1) * 2048 / R), 148);
  check_plan(cfg(1024, 4, 2, 256, 0, 256), 148);
  check_plan(cfg(4096, 4, 3, 256, 0, 1024), 148);
  check_plan(cfg(65536, 1, 2, 128, 0, 65536), 148);
  check_plan(cfg(16384, 4, 4, 256, 0, 4096), 148);
  check_plan(cfg(16384, 4, 4, 512, 0, 4096), 148);
  check_plan(cfg(16384, 4, 4, 512, 0, 4096, 0), 148);
  check_plan(cfg(4 * 37, 4, 2, 96, 0, 37), 148);
  check_plan(cfg(32, 4, 2, 33, 0, 8), 148);
  check_plan(cfg(4, 4, 2, 16, 0, 1), 148);
  check_plan(cfg(512, 32, 1, 64, 0, 16), 148);
  check_plan(cfg(3 * 40, 3, 2, 128, 0, 40), 148);          // sequence lengths that are not powers of two
  check_plan(cfg(6 * 24, 6, 2, 64, 0, 24), 148);
  check_plan(cfg(5 * 2048, 5, 3, 256, 0, 2048), 148);
  check_plan(cfg(2048, 4, 8, 128, 0, 512), 148);           // 8 modalities
  // documented limits
  Plan p;
  CHECK(build_plan(cfg(8190, 4, 2, 256, 0, 2047), p, 148) == FOCAL_ESHAPE, "B %% S");
  CHECK(build_plan(cfg(33 * 8, 33, 2, 256, 0, 8), p, 148) == FOCAL_ESHAPE, "S > 32");
  CHECK(build_plan(cfg(8192, 4, 2, 514, 0, 2048), p, 148) == FOCAL_ESHAPE, "D > 512");
  CHECK(build_plan(cfg(8192, 4, 9, 256, 0, 2048), p, 148) == FOCAL_ESHAPE, "M > 8");
  CHECK(build_plan(cfg(8192, 4, 2, 256, 5, 5), p, 148) == FOCAL_EINVAL, "empty shard");
  if (fails) {
    std::printf("%d failures\n", fails);
    return 1;
  }
  std::printf("plan ok\n");
  return 0;
}
