// CPU replay of the slot protocol of the cluster Gauss-Jordan inverse (isrsolver_b200/csrc/linalg.cu) with the kernel's
// own phase arithmetic (gj_phases.cuh).  For every N from 1 to 240 the pivot blocks are processed in order on the
// simulated CTAs of a 4-, 8- and 16-CTA cluster; every mbarrier is a counter of completed phases, and at every wait of the kernel the parity it would
// pass must be the parity of the phase that has just completed for that use:
//   empty[slot]  owner of block B >= 4 waits for the four credits of block B - 4
//   lfull[slot]  warps of the owner's CTA wait for the locally staged block
//   rfull[slot]  the other CTAs wait for the copied block -- and the expect_tx outstanding on that barrier must have
//                been posted for exactly this block (thread 0 posts the next remote block of the slot after each wait)
//   done[slot]   the forwarder waits for all warps of its CTA, then credits the owner of block B + 4
// A wrong count would make the real kernel hang or read a stale buffer.
#include <cstdio>
#include <cstdlib>

#include "gj_phases.cuh"

using namespace isr;

#define CHECK(cond)                                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      std::printf("FAILED %s (line %d): n=%d block=%d cta=%d\n", #cond, __LINE__, n, B, c); \
      std::exit(1);                                                                      \
    }                                                                                    \
  } while (0)

// C: CTAs of the cluster (the kernel is launched with 4, 8 or 16); RW: matrix rows per warp (4 or 8: a CTA owns whole warps)
template <int C, int RW>
static long replay() {
  long waits = 0;
  for (int n = 1; n <= 240; ++n) {
    const int R = RW * ((n + RW * C - 1) / (RW * C)), NB = (n + 3) / 4;
    GjPhases ph[C];
    int rfull[C][kGjSlots] = {}, lfull[C][kGjSlots] = {}, empty[C][kGjSlots] = {}, done[C][kGjSlots] = {};
    int credits[C][kGjSlots] = {}, expected[C][kGjSlots];
    int B = -1, c = 0;
    for (c = 0; c < C; ++c) {
      const int Bs = c * R / 4;
      ph[c] = GjPhases{NB, Bs, Bs + R / 4 < NB ? Bs + R / 4 : NB};
      for (int q = 0; q < kGjSlots; ++q)      // the kernel's initial expect_tx: first remote block of every slot
        expected[c][q] = (q < NB && !ph[c].owned(q)) ? q : ph[c].next_remote(q);
    }
    for (B = 0; B < NB; ++B) {
      const int q = B & (kGjSlots - 1), o = 4 * B / R;
      c = o;
      CHECK(o < C && ph[o].owned(B));
      for (int x = 0; x < C; ++x) CHECK((x == o) == ph[x].owned(B));           // exactly one owner
      if (B >= kGjSlots) {                                                      // owner: credits of block B - 4
        CHECK(empty[o][q] >= 1 && ph[o].empty_parity(B) == (uint32_t)((empty[o][q] - 1) & 1));
        ++waits;
      }
      ++lfull[o][q];                                                            // staged block published locally
      for (c = 0; c < C; ++c) {
        if (c == o) {
          CHECK(ph[c].lfull_parity(B) == (uint32_t)((lfull[c][q] - 1) & 1));
        } else {
          CHECK(expected[c][q] == B);                                           // the outstanding expect_tx is for this block
          ++rfull[c][q];                                                        // expect_tx arrival + all bytes: phase done
          CHECK(ph[c].rfull_parity(B) == (uint32_t)((rfull[c][q] - 1) & 1));
          expected[c][q] = ph[c].next_remote(B);                                // thread 0 re-arms the barrier
        }
        ++waits;
        ++done[c][q];                                                           // all warps of the CTA are done with the slot
        CHECK(GjPhases::done_parity(B) == (uint32_t)((done[c][q] - 1) & 1));
        if (B + kGjSlots < NB) {
          const int o2 = 4 * (B + kGjSlots) / R;
          if (++credits[o2][q] == C) { credits[o2][q] = 0; ++empty[o2][q]; }    // four credits complete a phase of empty
        }
      }
    }
    B = NB;
    for (c = 0; c < C; ++c)
      for (int q = 0; q < kGjSlots; ++q) CHECK(expected[c][q] == -1 && credits[c][q] == 0);   // nothing left armed
  }
  return waits;
}

int main() {
  const long waits = replay<4, 4>() + replay<8, 4>() + replay<16, 4>() + replay<4, 8>() + replay<8, 8>() + replay<16, 8>();
  std::printf("gauss-jordan slot protocol ok: %ld waits replayed\n", waits);
  return 0;
}
