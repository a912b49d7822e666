// split_host.cu -- the row membership of train_test_split(X, y, test_size=0.2, random_state=42)
// (stage_1_train_model.py:98-103 -> sklearn/model_selection/_split.py: ShuffleSplit._iter_indices:
//  `permutation = rng.permutation(n_samples); ind_test = permutation[:n_test]; ind_train = permutation[n_test:n_test+n_train]`)
// as a one-byte-per-row mask, bit for bit what numpy's legacy RandomState draws:
//     MT19937 seeded by init_genrand(seed)                         (numpy/random/_mt19937: mt19937_seed)
//     permutation(n) = arange(n) shuffled by Fisher-Yates, i = n-1 .. 1, j = random_interval(i)  (mtrand.pyx _shuffle_raw)
//     random_interval(max): mask = smallest 2^k - 1 >= max; draw 32-bit words (64-bit when max > 2^32 - 1), AND with the
//                           mask, reject values > max                (numpy/random/src/legacy/legacy-distributions.c)
// This is HOST logic by nature -- a sequential generator feeding a sequential shuffle -- exactly where the reference does
// it; what the library adds is speed at scale: the swap partner j depends only on the generator, so a window of draws is
// produced ahead of the shuffle and the cache lines of a[j] are prefetched (the shuffle of a multi-GB index array is
// otherwise one DRAM miss per row: numpy needs ~115 ns / row at 10^8 rows, this loop ~10x less), on a uint32 index
// array (half the traffic of numpy's int64).  The GPU consumes the mask (row_mask of b2_gram_accumulate / b2_score).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "b2_internal.cuh"

namespace {

struct Mt19937 {
  uint32_t mt[624];
  int pos;
  explicit Mt19937(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    pos = 624;
  }
  void refill() {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    pos = 0;
  }
  inline uint32_t next32() {
    if (pos >= 624) refill();
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline uint64_t next64() { const uint64_t hi = next32(); return (hi << 32) | next32(); }
  inline uint64_t interval(uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) { while ((v = (next32() & mask)) > max) {} }
    else { while ((v = (next64() & mask)) > max) {} }
    return v;
  }
};

template <typename I>
int shuffle_and_mark(int64_t n, int64_t n_test, uint32_t seed, uint8_t* mask) {
  I* a = static_cast<I*>(malloc((size_t)n * sizeof(I)));
  if (a == nullptr) { b2::set_error("out of host memory for the %lld-row permutation", (long long)n); return B2_E_STATE; }
  for (int64_t i = 0; i < n; ++i) a[i] = (I)i;
  Mt19937 rng(seed);
  constexpr int kAhead = 64;                       // draws produced (and prefetched) ahead of the swap that uses them
  int64_t jbuf[kAhead];
  int64_t produced = n - 1;                        // next i whose partner has not been drawn yet
  for (int k = 0; k < kAhead && produced >= 1; ++k, --produced) {
    jbuf[(n - 1 - produced) % kAhead] = (int64_t)rng.interval((uint64_t)produced);
    __builtin_prefetch(a + jbuf[(n - 1 - produced) % kAhead], 1, 0);
  }
  for (int64_t i = n - 1; i >= 1; --i) {
    const int slot = (int)((n - 1 - i) % kAhead);
    const int64_t j = jbuf[slot];
    const I t = a[j]; a[j] = a[i]; a[i] = t;
    if (produced >= 1) {                            // refill the slot just used with the partner of i - kAhead
      jbuf[slot] = (int64_t)rng.interval((uint64_t)produced);
      __builtin_prefetch(a + jbuf[slot], 1, 0);
      --produced;
    }
  }
  memset(mask, 1, (size_t)n);
  for (int64_t k = 0; k < n_test; ++k) {
    if (k + 32 < n_test) __builtin_prefetch(mask + a[k + 32], 1, 0);
    mask[a[k]] = 0;
  }
  free(a);
  return B2_OK;
}

}  // namespace

extern "C" int b2_split_mask(int64_t n_rows, int64_t n_test, uint32_t seed, uint8_t* mask_out) {
  if (n_rows < 1 || n_test < 0 || n_test > n_rows || mask_out == nullptr) {
    b2::set_error("b2_split_mask: need n_rows >= 1, 0 <= n_test <= n_rows and a mask buffer");
    return B2_E_ARG;
  }
  if (n_rows <= 0xffffffffll) return shuffle_and_mark<uint32_t>(n_rows, n_test, seed, mask_out);
  return shuffle_and_mark<int64_t>(n_rows, n_test, seed, mask_out);
}
