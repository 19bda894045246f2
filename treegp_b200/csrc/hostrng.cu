// Host-side bootstrap index generation, bit-identical to numpy.
//
// The reference draws every bootstrap resample with
//     rng = numpy.random.default_rng(seed);  rng.integers(0, npsfs - 1, size=npsfs)
// (/root/reference/treegp/two_pcf.py:266,269-281 -- the exclusive upper bound means index n-1 is never drawn).
// At N = 40 000 and 444 resamples numpy needs ~0.1 s for the 1.8e7 draws, more than all the pair counting they
// feed.  This is the same generator restated in C: PCG64 (XSL-RR 128/64, numpy's `PCG64`), its 32-bit output
// buffering (`has_uint32` / `uinteger`), and the bounded-integer algorithm numpy uses for ranges below 2^32
// (Lemire multiply-shift with rejection, `buffered_bounded_lemire_uint32`).  Instead of materialising the b x n
// index array (142 MB at the sizes above) the draws are turned into per-point multiplicities on the fly.
// Pinned against numpy itself in tests/test_cpu_host.py (streams and final generator state).
#include <stdint.h>
#include <string.h>
#include "tgp_common.cuh"

namespace {
typedef unsigned __int128 u128;
struct Pcg64 {
  u128 state, inc;
  int has_uint32;
  uint32_t uinteger;
};
inline uint64_t rotr64(uint64_t v, unsigned r) { return (v >> r) | (v << ((-r) & 63)); }
inline uint64_t next64(Pcg64& g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  g.state = g.state * mult + g.inc;
  const uint64_t hi = (uint64_t)(g.state >> 64), lo = (uint64_t)g.state;
  return rotr64(hi ^ lo, (unsigned)(hi >> 58));
}
inline uint32_t next32(Pcg64& g) {
  if (g.has_uint32) {
    g.has_uint32 = 0;
    return g.uinteger;
  }
  const uint64_t v = next64(g);
  g.has_uint32 = 1;
  g.uinteger = (uint32_t)(v >> 32);
  return (uint32_t)v;
}
// uniform integer in [0, rng] (inclusive), rng < 0xFFFFFFFF
inline uint32_t bounded_lemire32(Pcg64& g, uint32_t rng) {
  const uint32_t rng_excl = rng + 1u;
  uint64_t m = (uint64_t)next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)next32(g) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return (uint32_t)(m >> 32);
}
}  // namespace

extern "C" int tgp_bootstrap_multiplicities(uint64_t* state, int64_t n, int64_t b, const int64_t* pos,
                                            uint8_t* mult) {
  TGP_CHECK_ARG(state && mult, "null pointer");
  TGP_CHECK_ARG(n >= 2 && n - 2 < 0xFFFFFFFFll && b >= 0, "need 2 <= n <= 2^32 and b >= 0");
  Pcg64 g;
  g.state = ((u128)state[0] << 64) | state[1];
  g.inc = ((u128)state[2] << 64) | state[3];
  g.has_uint32 = state[4] != 0;
  g.uinteger = (uint32_t)state[5];
  const uint32_t rng = (uint32_t)(n - 2);   // integers(0, n-1): values 0 .. n-2
  memset(mult, 0, (size_t)(b * n));
  int overflow = 0;
  for (int64_t r = 0; r < b; ++r) {
    uint8_t* row = mult + r * n;
    if (rng == 0) {            // numpy draws nothing for a one-value range
      row[pos ? pos[0] : 0] = (uint8_t)(n > 255 ? 255 : n);
      overflow |= n > 255;
      continue;
    }
    const uint32_t rng_excl = rng + 1u;
    int64_t i = 0;
    while (i < n) {
      if (g.has_uint32 || i + 1 >= n) {          // odd position in the 64-bit stream, or last draw of the row
        const uint32_t idx = bounded_lemire32(g, rng);
        ++row[pos ? pos[idx] : idx];
        ++i;
        continue;
      }
      // two draws from one 64-bit output (low half first, as numpy's next_uint32 does)
      const uint64_t v = next64(g);
      const uint64_t m0 = (uint64_t)(uint32_t)v * rng_excl, m1 = (v >> 32) * rng_excl;
      if ((uint32_t)m0 < rng_excl || (uint32_t)m1 < rng_excl) {
        // (rare, ~n / 2^32 per draw) a value that may be rejected: hand both halves to the generic path
        g.has_uint32 = 1;
        g.uinteger = (uint32_t)(v >> 32);
        uint64_t m = m0;
        uint32_t leftover = (uint32_t)m;
        if (leftover < rng_excl) {
          const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
          while (leftover < threshold) {
            m = (uint64_t)next32(g) * rng_excl;
            leftover = (uint32_t)m;
          }
        }
        const uint32_t idx = (uint32_t)(m >> 32);
        ++row[pos ? pos[idx] : idx];
        ++i;
        continue;
      }
      g.uinteger = (uint32_t)(v >> 32);          // what numpy leaves in its (now consumed) 32-bit buffer
      const uint32_t i0 = (uint32_t)(m0 >> 32), i1 = (uint32_t)(m1 >> 32);
      ++row[pos ? pos[i0] : i0];
      ++row[pos ? pos[i1] : i1];
      i += 2;
    }
    // a uint8 counter that wrapped shows up as a short row sum
    int64_t total = 0;
    for (int64_t c = 0; c < n; ++c) total += row[c];
    overflow |= (total != n);
  }
  state[0] = (uint64_t)(g.state >> 64);
  state[1] = (uint64_t)g.state;
  state[4] = (uint64_t)g.has_uint32;
  state[5] = g.uinteger;
  if (overflow) {
    tgp_set_error("tgp_bootstrap_multiplicities: a point was drawn more than 255 times");
    return TGP_ERR_UNSUPPORTED;
  }
  return TGP_OK;
}
