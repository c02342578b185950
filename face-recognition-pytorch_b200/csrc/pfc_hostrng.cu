// Host side of the sampling draw (nets/PartialFC.py:110): `perm = torch.rand(size=[num_local])` on the CPU default
// generator, every step and rank.  The index set of the sampled classes is a pure function of that draw, so bit-exact
// sampling needs exactly torch's numbers -- and torch produces them one virtual call at a time (2-8 ns per float: 0.8 ms
// for a 360 k-class shard, 4.5 ms for 2 M classes on the bench box), which makes the HOST generator the bottleneck of a
// sampled step that takes 0.23 ms on the GPU (profiles/r02e_ref_gpu_cfg3.json).
//
// pfc_host_mt19937_uniform restates the published algorithm torch's CPU generator uses -- MT19937 (Matsumoto & Nishimura
// 1998: 624-word state, twist with 0x9908b0df, the four tempering shifts), float32 uniform = (word & (2^24 - 1)) * 2^-24
// (ATen's uniform_real transformation for a 24-bit mantissa) -- as a bulk generator working directly on the state blob
// `torch.Generator.get_state()` returns for a CPU generator (legacy layout: uint64 seed, int32 left, int32 seeded,
// uint64 next, uint64 state[624], then the cached-normal fields, 5056 bytes) and writes the advanced state back, so the
// generator continues exactly as if torch.rand had been called.  Plain host code: no CUDA call, no allocation.
// Pinned by tests/test_host_rng.py against torch.rand itself (sizes, state positions, interleaved torch calls).
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "pfc_internal.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr size_t TORCH_CPU_STATE_BYTES = 5056;
constexpr size_t OFF_LEFT = 8, OFF_SEEDED = 12, OFF_NEXT = 16, OFF_STATE = 24;

inline uint32_t twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

// the next 624 words, in place
void regenerate(uint32_t* s) {
    int k = 0;
    for (; k < MT_N - MT_M; ++k) s[k] = s[k + MT_M] ^ twist(s[k], s[k + 1]);
    for (; k < MT_N - 1; ++k) s[k] = s[k + MT_M - MT_N] ^ twist(s[k], s[k + 1]);
    s[MT_N - 1] = s[MT_M - 1] ^ twist(s[MT_N - 1], s[0]);
}

inline float temper_to_unit(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return static_cast<float>(y & 0xffffffu) * (1.0f / 16777216.0f);
}

}  // namespace

extern "C" {

size_t pfc_host_mt19937_state_bytes(void) { return TORCH_CPU_STATE_BYTES; }

// state: the 5056-byte blob of torch.Generator(device="cpu").get_state(), updated in place; out[n] <- the next n values
// torch.rand(n, dtype=float32) would return.
int pfc_host_mt19937_uniform(uint8_t* state, size_t state_bytes, float* out, size_t n) {
    if (!state || state_bytes != TORCH_CPU_STATE_BYTES || (!out && n)) return PFC_ERR_SHAPE;
    int32_t left, seeded;
    uint64_t next;
    memcpy(&left, state + OFF_LEFT, 4);
    memcpy(&seeded, state + OFF_SEEDED, 4);
    memcpy(&next, state + OFF_NEXT, 8);
    if (seeded != 1 || left < 1 || left > MT_N || next > static_cast<uint64_t>(MT_N)) return PFC_ERR_SHAPE;
    uint32_t s[MT_N];
    for (int k = 0; k < MT_N; ++k) {
        uint64_t w;
        memcpy(&w, state + OFF_STATE + 8 * static_cast<size_t>(k), 8);
        s[k] = static_cast<uint32_t>(w);
    }
    // torch's engine: `if (--left == 0) next_state();  y = state[next++]`; a freshly seeded engine has left = 1, next = 0
    size_t pos = (left == 1) ? MT_N : static_cast<size_t>(next);     // words of the current block already handed out
    if (left != 1 && pos != static_cast<size_t>(MT_N - left + 1)) return PFC_ERR_SHAPE;
    size_t i = 0;
    while (i < n) {
        if (pos == MT_N) {
            regenerate(s);
            pos = 0;
        }
        size_t take = MT_N - pos;
        if (take > n - i) take = n - i;
        const uint32_t* src = s + pos;
        float* dst = out + i;
        for (size_t k = 0; k < take; ++k) dst[k] = temper_to_unit(src[k]);
        pos += take;
        i += take;
    }
    if (n) {
        // pos in [1, 624] here: exactly the engine's state after its n-th output
        left = static_cast<int32_t>(MT_N - pos + 1);
        next = pos;
        memcpy(state + OFF_LEFT, &left, 4);
        memcpy(state + OFF_NEXT, &next, 8);
        for (int k = 0; k < MT_N; ++k) {
            const uint64_t w = s[k];
            memcpy(state + OFF_STATE + 8 * static_cast<size_t>(k), &w, 8);
        }
    }
    return PFC_OK;
}

}  // extern "C"
