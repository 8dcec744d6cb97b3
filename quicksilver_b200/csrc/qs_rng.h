/* qs_rng.h -- per-particle random number streams, shared by host C++ and device code.
 *
 * Follows the reference's MC_RNG_State (src/MC_RNG_State.hh:23-32, src/MC_RNG_State.cc:29-113):
 * a 64-bit LCG whose state doubles as the particle's seed, and a 2-round "pseudo-DES"
 * hash (Numerical Recipes psdes) that derives a child stream from a parent state.
 * Pure integer arithmetic + one u64->f64 conversion, so host and device agree bit for bit.
 */
#ifndef QS_RNG_H
#define QS_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define QS_RNG_HD __host__ __device__ __forceinline__
#else
#define QS_RNG_HD static inline
#endif

/* advance the stream, return a uniform double in (0,1)  (src/MC_RNG_State.hh:23-32) */
QS_RNG_HD double qs_rng_sample(uint64_t* state)
{
    *state = 2862933555777941757ULL * (*state) + 3037000493ULL;
    return 5.4210108624275222e-20 * (double)(*state);
}

/* 64 -> 64 bit hash, two Feistel rounds on the (high, low) words (src/MC_RNG_State.cc:29-52) */
QS_RNG_HD uint64_t qs_rng_hash(uint64_t v)
{
    uint32_t left = (uint32_t)(v >> 32), right = (uint32_t)v;
    const uint32_t ka[2] = { 0xbaa96887u, 0x1e17d32cu };
    const uint32_t kb[2] = { 0x4b0f3b58u, 0xe874f0c3u };
    for (int round = 0; round < 2; ++round)
    {
        uint32_t keep = right;
        uint32_t a  = right ^ ka[round];
        uint32_t lo = a & 0xffffu, hi = a >> 16;
        uint32_t b  = lo * lo + ~(hi * hi);
        uint32_t sw = (b >> 16) | ((b & 0xffffu) << 16);
        right = left ^ ((sw ^ kb[round]) + lo * hi);
        left  = keep;
    }
    return ((uint64_t)left << 32) | (uint64_t)right;
}

/* child seed = hash(parent state); the parent then advances once (src/MC_RNG_State.cc:107-113) */
QS_RNG_HD uint64_t qs_rng_spawn(uint64_t* parent)
{
    uint64_t child = qs_rng_hash(*parent);
    (void)qs_rng_sample(parent);
    return child;
}

#endif
