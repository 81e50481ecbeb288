// The device toy generator's bit source (isrsolver_b200/csrc/rng_device.cuh) on the CPU against the known-answer vectors
// of Philox4x32-10 published with Random123 (kat_vectors): counter / key -> output.
#include <cstdio>
#include <cstdlib>

#include "rng_device.cuh"

int main() {
  struct Kat { uint32_t c[4], k[2], out[4]; };
  const Kat kats[] = {
      {{0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}, {0x00000000u, 0x00000000u}, {0x6627e8d5u, 0xe169c58du, 0xbc57ac4cu, 0x9b00dbd8u}},
      {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}, {0x408f276du, 0x41c83b0eu, 0xa20bc7c6u, 0x6d5451fdu}},
      {{0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, {0xa4093822u, 0x299f31d0u}, {0xd16cfe09u, 0x94fdccebu, 0x5001e420u, 0x24126ea1u}},
  };
  for (const Kat& t : kats) {
    uint32_t r[4];
    isr::philox4x32_10(t.c[0], t.c[1], t.c[2], t.c[3], t.k[0], t.k[1], r);
    for (int q = 0; q < 4; ++q)
      if (r[q] != t.out[q]) {
        std::printf("FAILED: counter %08x.. word %d: %08x, expected %08x\n", t.c[0], q, r[q], t.out[q]);
        return 1;
      }
  }
  std::printf("philox4x32-10 known answers ok\n");
  return 0;
}
