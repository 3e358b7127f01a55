// prn_hostrng.cu — host side of the plane surface-normal term's triplet sampling (models/functions/vnl.py:48-53):
//   p = np.random.choice(n, k, replace=True); np.random.shuffle(p)      three times per region, numpy's GLOBAL legacy RNG
// restated in C so that a training step draws the SAME triplets as the reference for a given np.random.seed without spending
// ~30 ns per element inside numpy (64 ms per batch of 8 at 480x640 — twice the whole network's forward + backward).
// The legacy RandomState stream is frozen by numpy's compatibility policy (NEP 19):
//   choice(n, k)  -> randint(0, n, k)  -> masked rejection on 32-bit MT19937 outputs (no draw at all when n == 1)
//   shuffle(p)    -> for i = k-1 .. 1: j = random_interval(i) (masked rejection, 32-bit), swap(p[i], p[j])
// The caller passes numpy's MT19937 state in (np.random.get_state()) and puts the advanced state back (set_state), so the global
// stream continues exactly as if numpy had made the draws.  tests/test_plane_normal_cpu.py checks bit-equality with numpy.
// No device code: this file only rides in the same shared library.
#include <stdint.h>
#include <string.h>

#include "prn_internal.h"

namespace {

constexpr int kN = 624, kM = 397;

// MT19937 with the tempered outputs of a whole 624-word block produced at once (both loops vectorise), so that the consumers
// below are plain array reads.
struct Mt {
  uint32_t* key;
  int pos;
  uint32_t out[kN];
  bool have = false;
  inline void temper_all() {
    for (int i = 0; i < kN; ++i) {
      uint32_t y = key[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      out[i] = y;
    }
    have = true;
  }
  inline void gen() {
    int kk;
    uint32_t y;
    for (kk = 0; kk < kN - kM; ++kk) {
      y = (key[kk] & 0x80000000u) | (key[kk + 1] & 0x7fffffffu);
      key[kk] = key[kk + kM] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    for (; kk < kN - 1; ++kk) {
      y = (key[kk] & 0x80000000u) | (key[kk + 1] & 0x7fffffffu);
      key[kk] = key[kk + (kM - kN)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    y = (key[kN - 1] & 0x80000000u) | (key[0] & 0x7fffffffu);
    key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    pos = 0;
    temper_all();
  }
  // makes at least one tempered output available at out[pos]; returns how many are (pos .. kN)
  inline int avail() {
    if (pos == kN) gen();
    else if (!have) temper_all();
    return kN - pos;
  }
};

inline uint32_t mask_of(uint32_t v) {     // smallest 2^b - 1 >= v (v > 0)
  return 0xffffffffu >> __builtin_clz(v);
}

}  // namespace

extern "C" int prn_numpy_choice_shuffle(uint32_t* mt_key624, int32_t* mt_pos, const int64_t* n_of_region, const int64_t* k_of_region,
                                        int32_t n_regions, int32_t repeats, int32_t* out, int64_t out_stride) {
  using namespace prn;
  PRN_REQUIRE(mt_key624 && mt_pos && n_of_region && k_of_region && out && n_regions >= 0 && repeats >= 1 && *mt_pos >= 0 &&
                  *mt_pos <= kN, "numpy_choice_shuffle: bad arguments");
  Mt mt;
  mt.key = mt_key624;
  mt.pos = *mt_pos;
  int64_t off = 0;
  for (int r = 0; r < n_regions; ++r) {
    const int64_t n = n_of_region[r], k = k_of_region[r];
    PRN_REQUIRE(k >= 0 && (k == 0 || (n >= 1 && n <= 0xffffffffLL)) && off + k <= out_stride, "numpy_choice_shuffle: bad region %d", r);
    if (k == 0) continue;
    for (int j = 0; j < repeats; ++j) {
      int32_t* p = out + j * out_stride + off;
      const uint32_t rng = static_cast<uint32_t>(n - 1);
      if (rng == 0) {
        memset(p, 0, sizeof(int32_t) * k);
      } else {
        // masked rejection, branch-free: every draw is stored, the write cursor only advances past accepted ones (an
        // unpredictable accept/reject branch costs more than the generator itself)
        const uint32_t mask = mask_of(rng);
        int64_t i = 0;
        while (i < k) {
          const int n_av = mt.avail();
          const uint32_t* src = mt.out + mt.pos;
          int u = 0;
          // p has k slots and i < k inside the loop: the speculative store p[i] is always in bounds
          for (; u < n_av && i < k; ++u) {
            const uint32_t v = src[u] & mask;
            p[i] = static_cast<int32_t>(v);
            i += (v <= rng);
          }
          mt.pos += u;
        }
      }
      // Fisher-Yates from the top, j = random_interval(i): same trick, a rejected draw swaps p[i] with itself
      int64_t i = k - 1;
      while (i >= 1) {
        const int n_av = mt.avail();
        const uint32_t* src = mt.out + mt.pos;
        int u = 0;
        for (; u < n_av && i >= 1; ++u) {
          const uint32_t mx = static_cast<uint32_t>(i);
          const uint32_t v = src[u] & mask_of(mx);
          const bool ok = v <= mx;
          const uint32_t j = ok ? v : mx;
          const int32_t t = p[j];
          p[j] = p[i];
          p[i] = t;
          i -= ok;
        }
        mt.pos += u;
      }
    }
    off += k;
  }
  *mt_pos = mt.pos;
  return PRN_OK;
}
