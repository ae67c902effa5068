// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw, SC'11)
// and the samplers built on it.  Replaces the reference's global NumPy
// RandomState (MCMC.py:497 multivariate_normal, :331 rand, :160 gamma.rvs).
//
// Stream layout (results are independent of how chains are sharded over GPUs):
//     key     = (seed_lo, seed_hi)
//     counter = (chain_lo, chain_hi, iteration, slot)
//     slot 0,1   proposal normals (Box-Muller pairs)
//     slot 2     acceptance uniform
//     slot 4+2j  normal of gamma attempt j       slot 5+2j  uniform of gamma attempt j
// Doubles use 53 random bits: u = (k + 0.5) * 2^-53 in (0, 1).
#pragma once

#include <cstdint>

namespace rsfm {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ PhiloxKey philox_key(unsigned long long seed)
{
    PhiloxKey k;
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey k)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.k0, lo1, hi0 ^ c.w ^ k.k1, lo0);
        k.k0 += 0x9E3779B9u;
        k.k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 philox_block(PhiloxKey key, unsigned long long chain, uint32_t iter, uint32_t slot)
{
    return philox4x32_10(make_uint4((uint32_t)chain, (uint32_t)(chain >> 32), iter, slot), key);
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 11;
    return ((double)k + 0.5) * 0x1.0p-53;
}

__device__ __forceinline__ double philox_uniform(PhiloxKey key, unsigned long long chain, uint32_t iter, uint32_t slot)
{
    const uint4 r = philox_block(key, chain, iter, slot);
    return u53(r.x, r.y);
}

__device__ __forceinline__ void philox_normal2(PhiloxKey key, unsigned long long chain, uint32_t iter, uint32_t slot,
                                               double &z0, double &z1)
{
    const uint4 r = philox_block(key, chain, iter, slot);
    const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// Unit-scale Gamma(shape), shape > 1: Marsaglia & Tsang (2000), the algorithm
// behind NumPy's legacy standard_gamma that scipy.stats.gamma.rvs draws from.
// `tries` (optional, test hook): number of attempts consumed.
__device__ __forceinline__ double philox_gamma(PhiloxKey key, unsigned long long chain, uint32_t iter, double shape,
                                               uint32_t *tries = nullptr)
{
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t j = 0; j < 64; j++) {
        double x, unused;
        philox_normal2(key, chain, iter, 4u + 2u * j, x, unused);
        double v = 1.0 + c * x;
        if (tries) *tries = j + 1u;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = philox_uniform(key, chain, iter, 5u + 2u * j);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
    }
    return d;   // unreachable in practice (acceptance > 0.95 per attempt)
}

}  // namespace rsfm
