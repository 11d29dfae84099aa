// tg_rng.cuh - the reference's random stream ON THE DEVICE: MT19937 as numpy's legacy RandomState drives it.
//
// The reference draws every reset's randomness from `self.np_random` (rl_envs/base_tactile_env.py:61-64: gym <= 0.21 seeding ->
// a numpy RandomState) with uniform / randint / choice / rand (edge_follow_env.py:240,293; object_balance_env.py:300-313,
// 366-371; base_surface_env.py:290-309,448,508-514; object_push_env.py:204-229,289,310; object_roll_env.py:182-250).  The host can
// produce those draws and stream them through a ring (tg_set_draws / tg_draws_upload); here each env instead carries its own
// MT19937 state [624 words + position] in HBM and the reset code pulls from it with the SAME call semantics, so an env seeded like
// the reference consumes exactly the reference's sequence with no host in the loop:
//   random_sample  = (a >> 5) * 2^26 + (b >> 6)) / 2^53 from two outputs (rk_double)
//   uniform(lo,hi) = lo + (hi - lo) * random_sample
//   randint(0, n)  = masked rejection on 32-bit outputs (mask = next power of two - 1; `buffered_bounded_masked_uint32`)
//   choice([-1,1]) = randint(0, 2) -> one output & 1
// SURVEY.md section 7, hard part 4 ("ship a counter-based device RNG beside external draws"): this is that generator, chosen to
// be the reference's own rather than a new counter-based one so that seeds keep their meaning.
#pragma once
#include <stdint.h>

#define MT_N 624
#define MT_M 397

// draw kinds of TgTask.draw_kind
#define TG_DRAW_CONST 0          // draw_default[d], consumes nothing
#define TG_DRAW_UNIFORM 1        // uniform(draw_lo, draw_hi)
#define TG_DRAW_RANDINT 2        // randint(draw_hi) as a double
#define TG_DRAW_CHOICE_PM1 3     // choice([-1, 1])
#define TG_DRAW_CHOICE_RAND 4    // choice([-1, 1]) * rand()

struct MtState {
    uint32_t* key; // [624] this env's state words (global memory)
    int pos;       // next word; 624 = regenerate first (numpy's RandomState after seeding)
};

__device__ __noinline__ void mt_regenerate(uint32_t* key)
{
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    int kk = 0;
    for (; kk < MT_N - MT_M; kk++) {
        const uint32_t y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < MT_N - 1; kk++) {
        const uint32_t y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
        key[kk] = key[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    const uint32_t y = (key[MT_N - 1] & UPPER) | (key[0] & LOWER);
    key[MT_N - 1] = key[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

__device__ __forceinline__ uint32_t mt_next(MtState& s)
{
    if (s.pos >= MT_N) { mt_regenerate(s.key); s.pos = 0; }
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ double mt_double(MtState& s)
{
    const uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

// lo + (hi - lo) * u with TWO roundings like numpy's C (no fused multiply-add: a contracted FMA differs in the last bit, and the
// reset's IK turns one ulp of a draw into ~1e-7 rad of start pose)
__device__ __forceinline__ double mt_uniform(MtState& s, double lo, double hi)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(lo, __dmul_rn(hi - lo, mt_double(s)));
#else
    const volatile double prod = (hi - lo) * mt_double(s);
    return lo + prod;
#endif
}

// randint(0, n), 0 < n <= 2^32
__device__ __forceinline__ uint32_t mt_randint(MtState& s, double n)
{
    const uint32_t rng = (uint32_t)(n - 1.0);
    if (rng == 0u) return 0u;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = mt_next(s) & mask; } while (v > rng);
    return v;
}

__device__ __forceinline__ double mt_draw(MtState& s, int kind, double lo, double hi, double dflt)
{
    if (kind == TG_DRAW_UNIFORM) return mt_uniform(s, lo, hi);
    if (kind == TG_DRAW_RANDINT) return (double)mt_randint(s, hi);
    if (kind == TG_DRAW_CHOICE_PM1) return mt_randint(s, 2.0) ? 1.0 : -1.0;
    if (kind == TG_DRAW_CHOICE_RAND) {
        const double sg = mt_randint(s, 2.0) ? 1.0 : -1.0;
        return sg * mt_double(s);
    }
    return dflt;
}
