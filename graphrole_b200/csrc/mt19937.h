// MT19937 with NumPy's legacy RandomState conventions (host only).
//
// The reference's quantiser is KMeans(..., random_state=1) (graphrole/roles/factor.py:41): scikit-
// learn seeds np.random.RandomState(1) and draws (a) one `choice(n, p=uniform)` for the first
// k-means++ centre and (b) `uniform(size=trials)` per further centre (sklearn/cluster/_kmeans.py
// :231, :249).  Reproducing the reference's quantisation therefore needs the same stream: the
// public MT19937 recurrence (Matsumoto & Nishimura 1998), init_genrand seeding for integer seeds
// and the 53-bit double construction (a >> 5, b >> 6) NumPy's random_sample uses.
#pragma once

#include <stdint.h>

namespace gr {

class NumpyRandomState {
  public:
    explicit NumpyRandomState(uint32_t seed) {
        for (int i = 0; i < kN; ++i) {
            key_[i] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
        }
        pos_ = kN;
    }

    uint32_t next_u32() {
        if (pos_ == kN) refill();
        uint32_t y = key_[pos_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // RandomState.random_sample(): uniform on [0, 1) with 53 random bits
    double random_sample() {
        const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }

  private:
    static constexpr int kN = 624, kM = 397;
    uint32_t key_[kN];
    int pos_;

    void refill() {
        for (int i = 0; i < kN; ++i) {
            const uint32_t y = (key_[i] & 0x80000000u) | (key_[(i + 1) % kN] & 0x7fffffffu);
            key_[i] = key_[(i + kM) % kN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        pos_ = 0;
    }
};

}  // namespace gr
