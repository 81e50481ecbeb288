// CPU walk of the toy kernel's chunk schedules (isrsolver_b200/csrc/toy_gemm.cuh: SchedDense, SchedTri), the logic that
// decides which TMA loads are issued in which order.  No GPU is needed: the structs are __host__ __device__ and this
// file is compiled by nvcc and run on the host.  For every tile count, column-block count, chunk count, pipeline depth,
// pass selection and triangularity it checks that
//   (1) every (tile, pass, column block) is visited exactly once, its chunks kc = 0 .. count-1 in order, column blocks
//       in order, and a triangular column block b has exactly ceil(min(ld, (b + 1) NP) / KC) chunks;
//   (2) pass 2 of a tile starts at least STAGES - 1 chunks after the last chunk of its pass 1 (the d-tile is complete
//       before the first loads of its pass 2 are issued, loads running STAGES chunks ahead);
//   (3) NOP chunks appear only in the sequential order and the walk terminates after exactly the expected number of chunks;
//   (4) the triangular schedule (tri_first_live_tile, nk_tri) skips exactly the column tiles / chunks that hold nothing but
//       structural zeros of a lower-triangular operator, for every shipped tile shape.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

#include "toy_gemm.cuh"

using namespace isr;

static long long g_cases = 0;
#define CHECK(cond)                                                                                       \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      std::printf("FAILED %s at line %d: m=%d nblk=%d nk=%d stages=%d do2=%d tri=%d ld=%d\n", #cond, __LINE__, m, nblk, nk, \
                  stages, do2, tri, ld);                                                                  \
      std::exit(1);                                                                                       \
    }                                                                                                     \
  } while (0)

template <class S, int NP, int KC>
static void walk(int m, int nblk, int ld, int stages, int do2, int tri) {
  const int nk = (ld + KC - 1) / KC;
  S q;
  q.init(m, nblk, nk, stages - 1, do2, tri, ld);
  auto chunks_of = [&](int b, int pass2) {
    if (!((tri >> pass2) & 1)) return nk;
    const int lim = ld < (b + 1) * NP ? ld : (b + 1) * NP;
    return (lim + KC - 1) / KC;
  };
  std::map<std::tuple<int, int, int>, int> seen;          // (tile, pass, cb) -> chunks seen so far
  std::vector<long long> p1_end(m, -1), p2_begin(m, -1);
  long long pos = 0, expect = 0;
  int prev_tile[2] = {-1, -1};
  while (!q.done()) {
    const int kind = q.kind();
    CHECK(pos < 100000000LL);
    if (kind == PH_NOP) {
      CHECK(q.seq && do2);
    } else {
      const int pass = kind == PH_P2, t = q.tile();
      CHECK(t >= 0 && t < m && q.cb >= 0 && q.cb < nblk);
      CHECK(pass == 0 || do2);
      int& c = seen[{t, pass, q.cb}];
      CHECK(q.kc == c);                                   // chunks in order
      if (q.cb > 0 && c == 0) CHECK((seen[{t, pass, q.cb - 1}] == chunks_of(q.cb - 1, pass)));   // column blocks in order
      ++c;
      CHECK(c <= chunks_of(q.cb, pass));
      CHECK(q.last_chunk() == (c == chunks_of(q.cb, pass)));
      if (q.cb == 0 && q.kc == 0) { CHECK(t == prev_tile[pass] + 1); prev_tile[pass] = t; }     // tiles in order per pass
      if (pass == 0) p1_end[t] = pos;
      if (pass == 1 && p2_begin[t] < 0) p2_begin[t] = pos;
    }
    q.advance();
    ++pos;
  }
  for (int t = 0; t < m; ++t)
    for (int pass = 0; pass <= (do2 ? 1 : 0); ++pass)
      for (int b = 0; b < nblk; ++b) {
        CHECK((seen[{t, pass, b}] == chunks_of(b, pass)));
        expect += chunks_of(b, pass);
      }
  if (do2 && q.seq) expect += (long long)m * (stages - 1);
  CHECK(pos == expect);
  if (do2)
    for (int t = 0; t < m; ++t) CHECK(p2_begin[t] - p1_end[t] - 1 >= stages - 1);   // chunks strictly between the two
  ++g_cases;
}

template <int NP, int KC>
static void sweep() {
  for (int stages = 2; stages <= 4; ++stages)
    for (int m = 0; m <= 5; ++m)
      for (int nblk = 1; nblk <= 4; ++nblk)
        for (int do2 = 0; do2 <= 1; ++do2) {
          // ld: inside the last column block, at chunk boundaries and just off them
          const int lo = (nblk - 1) * NP + 2, hi = nblk * NP;
          const int lds[] = {lo, lo + KC - 2, (lo + hi) / 2 & ~1, hi - KC, hi - 2, hi};
          for (int ld : lds) {
            if (ld < lo || ld > hi || (ld & 1)) continue;
            walk<SchedDense, NP, KC>(m, nblk, ld, stages, do2, 0);
            for (int tri = 0; tri <= 3; ++tri) walk<SchedTri<NP, KC>, NP, KC>(m, nblk, ld, stages, do2, tri);
          }
        }
}

// Triangular skipping: with B[n][k] == 0 for k > n, (a) the chunks a column block does not load (kc >= nk_tri) hold only
// zeros for it, and (b) within a loaded chunk warp wn skips exactly the column tiles that are all zero there.
template <int WN, int BT, int KC>
static void tri_tiles(int nblk) {
  constexpr int NP = WN * BT * 8;
  const int m = 1, nk = 0, stages = 2, do2 = 1, tri = 3;   // (context of the CHECK message)
  for (int ld = (nblk - 1) * NP + 2; ld <= nblk * NP; ld += 2) {
    const int nk_all = (ld + KC - 1) / KC;
    SchedTri<NP, KC> q;
    q.init(1, nblk, nk_all, 1, 1, 3, ld);
    for (int cb = 0; cb < nblk; ++cb) {
      const int loaded = q.nk_tri(cb, 0);
      for (int kc = 0; kc < nk_all; ++kc) {
        const int k0 = kc * KC, n_first = cb * NP, n_last_blk = (cb + 1) * NP - 1 < ld - 1 ? (cb + 1) * NP - 1 : ld - 1;
        if (kc >= loaded) { CHECK(k0 > n_last_blk); continue; }           // (a): every k of the chunk is beyond every column
        for (int wn = 0; wn < WN; ++wn) {
          const int jm = tri_first_live_tile<WN>(k0 - n_first, wn);
          for (int j = 0; j < BT; ++j) {
            const int n_last = n_first + (j * WN + wn) * 8 + 7;           // last column of the tile
            if (j < jm) CHECK(n_last < k0);                               // skipped => all zero in this chunk
            else CHECK(n_last >= k0);                                     // live => it has a k <= n (nothing more could go)
          }
        }
        ++g_cases;
      }
    }
  }
}

int main() {
  tri_tiles<5, 5, 40>(1); tri_tiles<5, 5, 40>(3);      // m64n200w20
  tri_tiles<4, 8, 40>(1); tri_tiles<4, 8, 40>(2);      // m64n256w16 (N = 500: two column blocks)
  tri_tiles<4, 4, 24>(4); tri_tiles<4, 4, 40>(4);      // m64n128w16
  tri_tiles<1, 13, 40>(2);                             // m128n104w16
  tri_tiles<2, 7, 40>(2); tri_tiles<2, 13, 24>(1);     // m128n112w16, m64n208w8
  sweep<200, 40>();
  sweep<256, 40>();
  sweep<128, 24>();
  sweep<104, 40>();
  sweep<64, 24>();
  sweep<8, 24>();     // a column block narrower than a chunk: units shorter than the pipeline -> sequential order
  std::printf("schedule walk ok: %lld cases\n", g_cases);
  return 0;
}
