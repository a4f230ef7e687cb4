/* rsvd_b200_rng.h — counter-based Gaussian test-matrix generator (Philox4x32-10 + Box-Muller).
 *
 * Replaces the reference's Omega source, initialize_random_matrix()
 * (multi_core_mkl_code/matrix_vector_functions_intel_mkl.c:458-486): float32 Gaussians from a
 * time(NULL)-seeded VSL MCG31 stream, widened to double and stored in linear (column-major) order.
 * Here entry number `i` of that linear order is a pure function of (seed, i), so
 *   - the sketch kernel can produce Omega tiles on the fly (Omega is never stored),
 *   - every GPU of a row partition sees the same Omega without communication,
 *   - the CPU oracle (oracle/shim/vsl_shim.c) reproduces it bit for bit.
 *
 * Bit reproducibility CPU <-> GPU: only integer ops, exact int->float conversions, IEEE sqrt and
 * fused multiply-adds are used (every product/sum is spelled as one fmaf, so there is no
 * compiler-dependent contraction).  Values are float32 like the reference's `float *r` buffer.
 *
 * Usable from C (gcc), C++ and CUDA.
 */
#ifndef RSVD_B200_RNG_H
#define RSVD_B200_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RSVD_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define RSVD_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define RSVD_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define RSVD_SQRT(a) __fsqrt_rn((a))
#define RSVD_MULHI(a, b) __umulhi((a), (b))
#define RSVD_F2I(f) __float_as_int((f))
#define RSVD_I2F(i) __int_as_float((i))
#else
#define RSVD_FMA(a, b, c) fmaf((a), (b), (c))
#define RSVD_SQRT(a) sqrtf((a))
#define RSVD_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
RSVD_HD int32_t rsvd_f2i_(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
RSVD_HD float rsvd_i2f_(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
#define RSVD_F2I(f) rsvd_f2i_((f))
#define RSVD_I2F(i) rsvd_i2f_((i))
#endif

#define RSVD_PHILOX_M0 0xD2511F53u
#define RSVD_PHILOX_M1 0xCD9E8D57u
#define RSVD_PHILOX_W0 0x9E3779B9u
#define RSVD_PHILOX_W1 0xBB67AE85u

/* Philox4x32-10 (Salmon et al., SC'11).  ctr[4] in/out, key[2]. */
RSVD_HD void rsvd_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = RSVD_MULHI(RSVD_PHILOX_M0, c[0]);
        uint32_t lo0 = RSVD_PHILOX_M0 * c[0];
        uint32_t hi1 = RSVD_MULHI(RSVD_PHILOX_M1, c[2]);
        uint32_t lo1 = RSVD_PHILOX_M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += RSVD_PHILOX_W0; k1 += RSVD_PHILOX_W1;
    }
}

/* ln(x) for x in (0,1), Cephes-style polynomial, FMA-only (reproducible, ~1 ulp). */
RSVD_HD float rsvd_logf(float x) {
    int32_t ix = RSVD_F2I(x);
    int32_t e = ((ix >> 23) & 0xff) - 126;
    float m = RSVD_I2F((ix & 0x007fffff) | 0x3f000000); /* [0.5,1) */
    /* branch-free (selects only): data-dependent branches would diverge inside the GEMM's producer warps */
    const int lt = m < 0.70710678118654752440f;
    e -= lt;
    float f = RSVD_FMA(m, lt ? 2.0f : 1.0f, -1.0f);
    float z = RSVD_FMA(f, f, 0.0f);
    float y = 7.0376836292E-2f;
    y = RSVD_FMA(y, f, -1.1514610310E-1f);
    y = RSVD_FMA(y, f, 1.1676998740E-1f);
    y = RSVD_FMA(y, f, -1.2420140846E-1f);
    y = RSVD_FMA(y, f, 1.4249322787E-1f);
    y = RSVD_FMA(y, f, -1.6668057665E-1f);
    y = RSVD_FMA(y, f, 2.0000714765E-1f);
    y = RSVD_FMA(y, f, -2.4999993993E-1f);
    y = RSVD_FMA(y, f, 3.3333331174E-1f);
    y = RSVD_FMA(y, f, 0.0f);
    y = RSVD_FMA(y, z, 0.0f);
    float fe = (float)e;
    y = RSVD_FMA(fe, -2.12194440e-4f, y);
    y = RSVD_FMA(-0.5f, z, y);
    float r = RSVD_FMA(f, 1.0f, y);
    r = RSVD_FMA(fe, 0.693359375f, r);
    return r;
}

/* (cos, sin) of 2*pi*u, u = t24 / 2^24, by octant reduction + Cephes polynomials, FMA-only. */
RSVD_HD void rsvd_sincos2pi(uint32_t t24, float *cs, float *sn) {
    uint32_t o = t24 >> 21;                                  /* octant 0..7 */
    float f = (float)(t24 & 0x1fffffu) * (1.0f / 2097152.0f); /* [0,1), exact */
    const int odd = (int)(o & 1u);
    f = odd ? RSVD_FMA(f, -1.0f, 1.0f) : f;                  /* exact */
    float a = RSVD_FMA(f, 0.78539816339744830962f, 0.0f);    /* [0, pi/4] */
    float z = RSVD_FMA(a, a, 0.0f);
    float ps = -1.9515295891E-4f;
    ps = RSVD_FMA(ps, z, 8.3321608736E-3f);
    ps = RSVD_FMA(ps, z, -1.6666654611E-1f);
    ps = RSVD_FMA(ps, z, 0.0f);
    float s = RSVD_FMA(ps, a, a);
    float pc = 2.443315711809948E-005f;
    pc = RSVD_FMA(pc, z, -1.388731625493765E-003f);
    pc = RSVD_FMA(pc, z, 4.166664568298827E-002f);
    float zz = RSVD_FMA(z, z, 0.0f);
    float c = RSVD_FMA(-0.5f, z, 1.0f);
    c = RSVD_FMA(pc, zz, c);
    /* theta = q*pi/2 +/- a  (+ for even octant, - for odd) */
    const uint32_t q = ((o + 1u) >> 1) & 3u;
    const float ss = odd ? -s : s;
    /* q: 0 -> (ss, c), 1 -> (c, -ss), 2 -> (-ss, -c), 3 -> (-c, ss)  as (sin, cos); selects only */
    const float sa = (q & 1u) ? c : ss;          /* |sin| source */
    const float ca = (q & 1u) ? ss : c;          /* |cos| source */
    const float sv = (q & 2u) ? -sa : sa;        /* sin is negated in quadrants 2, 3 */
    const float cv = (q == 1u || q == 2u) ? -ca : ca;   /* cos is negated in quadrants 1, 2 */
    *cs = cv; *sn = sv;
}

/* Two N(0,1) float samples from two 32-bit words. */
RSVD_HD void rsvd_box_muller(uint32_t xa, uint32_t xb, float *z0, float *z1) {
    float u1 = (float)(2u * (xa >> 9) + 1u) * (1.0f / 16777216.0f); /* (0,1), exact */
    float r2 = RSVD_FMA(rsvd_logf(u1), -2.0f, 0.0f);
    float rad = RSVD_SQRT(r2);
    float c, s;
    rsvd_sincos2pi(xb >> 8, &c, &s);
    *z0 = RSVD_FMA(rad, c, 0.0f);
    *z1 = RSVD_FMA(rad, s, 0.0f);
}

/* The four normals of counter block `blk` (linear entries 4*blk .. 4*blk+3). */
RSVD_HD void rsvd_normal4(uint64_t seed, uint64_t blk, float z[4]) {
    uint32_t c[4];
    c[0] = (uint32_t)blk; c[1] = (uint32_t)(blk >> 32); c[2] = 0u; c[3] = 0u;
    rsvd_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    rsvd_box_muller(c[0], c[1], &z[0], &z[1]);
    rsvd_box_muller(c[2], c[3], &z[2], &z[3]);
}

/* Entry `i` of the linear stream. */
RSVD_HD float rsvd_normal_at(uint64_t seed, uint64_t i) {
    float z[4];
    rsvd_normal4(seed, i >> 2, z);
    return z[i & 3u];
}

#endif /* RSVD_B200_RNG_H */
