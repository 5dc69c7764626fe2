// ssi_rng.cuh — counter-based RNG shared by device code and the host replay
// (ssi_rng_replay).  The reference draws from Julia's unseeded GLOBAL_RNG
// (src/space_inference.jl:113-116); this stream replaces it so that the host / the
// oracle can replay exactly what every chain saw.
//
//   Philox4x32-10, key = (seed_lo, seed_hi), counter = (chain, step, block, stream)
//   stream 0: normals.  block j -> x0..x3; u_i = (x_i + 0.5) 2^-32 (double);
//             (n0,n1) = sqrt(-2 ln u0) (cos,sin)(2 pi u1), (n2,n3) likewise from (u2,u3);
//             evaluated in double, rounded once to float.  Normal i lives in block i/4.
//   stream 1: accept variate.  block 0, e = -ln((x0 + 0.5) 2^-32)  ~ Exp(1), double.
#pragma once
#include <cstdint>
#include <cmath>

#ifdef __CUDACC__
#define SSI_HD __host__ __device__ __forceinline__
#else
#define SSI_HD inline
#endif

#define SSI_STREAM_NORMAL 0u
#define SSI_STREAM_ACCEPT 1u

struct ssi_u4 { uint32_t x, y, z, w; };

SSI_HD void ssi_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

SSI_HD ssi_u4 ssi_philox4x32_10(ssi_u4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        ssi_mulhilo(0xD2511F53u, c.x, hi0, lo0);
        ssi_mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
        ssi_u4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

SSI_HD double ssi_u01(uint32_t x) { return ((double)x + 0.5) * 2.3283064365386963e-10; /* 2^-32 */ }

// four N(0,1) floats of (chain, step, block)
SSI_HD void ssi_normal4(uint64_t seed, uint32_t chain, uint32_t step, uint32_t block, float out[4]) {
    ssi_u4 c{chain, step, block, SSI_STREAM_NORMAL};
    const ssi_u4 r = ssi_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double two_pi = 6.283185307179586476925286766559;
    {
        const double rad = sqrt(-2.0 * log(ssi_u01(r.x)));
        const double ang = two_pi * ssi_u01(r.y);
        out[0] = (float)(rad * cos(ang));
        out[1] = (float)(rad * sin(ang));
    }
    {
        const double rad = sqrt(-2.0 * log(ssi_u01(r.z)));
        const double ang = two_pi * ssi_u01(r.w);
        out[2] = (float)(rad * cos(ang));
        out[3] = (float)(rad * sin(ang));
    }
}

SSI_HD double ssi_exp1(uint64_t seed, uint32_t chain, uint32_t step) {
    ssi_u4 c{chain, step, 0u, SSI_STREAM_ACCEPT};
    const ssi_u4 r = ssi_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return -log(ssi_u01(r.x));
}
