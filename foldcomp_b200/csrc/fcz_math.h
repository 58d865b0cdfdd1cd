// foldcomp_b200/csrc/fcz_math.h -- per-item arithmetic of the FCZ hot path (host + device).
//
// Everything here is FCZ_HD so that the CUDA kernels (fcz_engine.cu) and the single-thread host
// model used by the CPU tests (tests/emu/) compile the SAME arithmetic.  Compile with FMA
// contraction disabled (nvcc -fmad=false, g++ -ffp-contract=off): the encode side must reproduce
// the reference's x86-64 float/double rounding sequence bit for bit.
#ifndef FCZ_MATH_H
#define FCZ_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FCZ_HD __host__ __device__ __forceinline__
#define FCZ_HD_SLOW __host__ __device__ __noinline__  // rare exact fallbacks: keep them out of the hot instruction stream
#else
#define FCZ_HD inline
#define FCZ_HD_SLOW inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace fcz {

struct f3 {
    float x, y, z;
};

FCZ_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
// Approximate reciprocal square root / reciprocal: MUFU.RSQ / MUFU.RCP on the device (rsqrt.approx / rcp.approx: <= 2 ulp, CUDA C
// Programming Guide, mathematical functions appendix), IEEE operations on the host.  Every use sits behind an error
// bound that covers several ulp either way.  FCZ_EMU_PERTURB (tests only, tests/emu) pushes the host results off by
// up to +-3 ulp, pseudo-randomly per argument, so that the CPU tests exercise those bounds and not just the
// correctly rounded special case.
#if !defined(__CUDA_ARCH__) && defined(FCZ_EMU_PERTURB)
inline float emu_perturb_(float v, float x) {
    uint32_t h, u;
    memcpy(&h, &x, 4);
    h = (h ^ (h >> 15)) * 0x2c1b3c6du; h ^= h >> 12;
    const int k = (int)(h % 7u) - 3;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) != 0x7f800000u && (u & 0x7fffffffu) > 8u) u = (uint32_t)((int32_t)u + k);
    memcpy(&v, &u, 4);
    return v;
}
#endif
FCZ_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
    float r;  // one MUFU.RSQ: rsqrtf() wraps it in denormal scaling that no caller needs (arguments are >= 1e-30 or rejected)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#elif defined(FCZ_EMU_PERTURB)
    return emu_perturb_(1.0f / sqrtf(x), x);
#else
    return 1.0f / sqrtf(x);
#endif
}
FCZ_HD float rcp_(float x) {
#if defined(__CUDA_ARCH__)
    float r;  // one MUFU.RCP
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#elif defined(FCZ_EMU_PERTURB)
    return emu_perturb_(1.0f / x, x);
#else
    return 1.0f / x;
#endif
}

FCZ_HD f3 mk3(float x, float y, float z) {
    f3 r;
    r.x = x; r.y = y; r.z = z;
    return r;
}
FCZ_HD f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
FCZ_HD void st3(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
FCZ_HD f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }

// reference: crossProduct, src/float3d.h:19-25 (float mul/sub, no FMA)
FCZ_HD f3 cross3(f3 a, f3 b) {
    return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// reference: norm, src/float3d.h:33-35 -- pow(float,2) promotes to double: exact squares, double
// adds, double sqrt, rounded to float on return.
FCZ_HD float norm3(f3 v) {
    double s = (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z;
    return (float)sqrt(s);
}

// reference: getCosineTheta, src/float3d.h:36-43 -- float dot/sizes, float product of sizes,
// DOUBLE sqrt and DOUBLE divide, rounded to float.
FCZ_HD float cos_theta(f3 v1, f3 v2) {
    float inner = (v1.x * v2.x) + (v1.y * v2.y) + (v1.z * v2.z);
    float s1 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    float s2 = v2.x * v2.x + v2.y * v2.y + v2.z * v2.z;
    return (float)((double)inner / sqrt((double)(s1 * s2)));
}

// ---- the same quantities, split so that the expensive double-precision steps can be CERTIFIED
// shortcuts: a cheaper evaluation is used only when it provably rounds to the same float as the
// reference's sequence; otherwise the reference's exact sequence runs.  Results are identical always.
struct DotParts {
    float inner, p;  // inner product and float product of the squared sizes, exactly as getCosineTheta forms them
};
FCZ_HD DotParts dot_parts(f3 v1, f3 v2) {
    DotParts d;
    d.inner = (v1.x * v2.x) + (v1.y * v2.y) + (v1.z * v2.z);
    float s1 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    float s2 = v2.x * v2.x + v2.y * v2.y + v2.z * v2.z;
    d.p = s1 * s2;
    return d;
}
FCZ_HD float cos_exact(DotParts d) { return (float)((double)d.inner / sqrt((double)d.p)); }
// the same, as a real call: used where the exact sequence is a rare fallback (otherwise the compiler
// speculates its first ~40 instructions into every iteration)
static FCZ_HD_SLOW float cos_exact_slow(float inner, float p) { return (float)((double)inner / sqrt((double)p)); }
static FCZ_HD_SLOW float deg_exact_slow(double ac) { return (float)(ac * 180.0 / M_PI); }

FCZ_HD double drsqrt_(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// Both roundings of an interval [v(1-2^-46), v(1+2^-46)] agree  =>  every double in it rounds to that float.
FCZ_HD bool same_float(double v, float* f) {
    const float lo = (float)(v * (1.0 - 0x1p-46)), hi = (float)(v * (1.0 + 0x1p-46));
    *f = lo;
    return lo == hi;  // false for NaN
}
// cos(theta) as the reference rounds it.  Fast path: inner * rsqrt(p) in double is within 2^-50 (relative)
// of the reference's RN(inner / RN(sqrt(p))) -- rsqrt <= 1 ulp, one multiply -- so if the whole 2^-46
// interval around it rounds to one float, that float is the reference's.  Fails with probability ~2^-21.
FCZ_HD float cos_ref(DotParts d) {
    float c;
    if (d.p >= 1e-30f && d.p <= 1e30f && same_float((double)d.inner * drsqrt_((double)d.p), &c)) return c;
    return cos_exact_slow(d.inner, d.p);
}
// degrees of an arccosine as the reference rounds them: (float)(ac * 180.0 / M_PI) is within 2^-50.5 of
// ac * RN(180/pi); certified the same way.
FCZ_HD float deg_ref(double ac) {
    float g;
    if (same_float(ac * 57.29577951308232, &g)) return g;
    return deg_exact_slow(ac);
}

// acos(c) in DEGREES for a float c, |relative error| <= 2^-46 over every float in [-1, 1] (verified
// exhaustively on the host, tests/emu: the function uses only IEEE operations -- fma, *, +, sqrt -- in a fixed
// order, so the device computes bit-identical values); NaN for |c| > 1 or NaN.  ~30 double instructions
// against ~105 for the library acos.  asin(s) = s + s z P(z), z = s^2 <= 1/4, P a degree-10 fit.
FCZ_HD double acos_deg_fast(float c) {
    const double x = (double)c, ax = fabs(x);
    const bool big = ax > 0.5;
    const double z = big ? (1.0 - ax) * 0.5 : x * x;  // exact for a float c
    const double s = big ? sqrt(z) : x;
    const double C[11] = {
        0x1.55555555555c8p-3,
        0x1.33333332ff5fap-4,
        0x1.6db6dbad27a33p-5,
        0x1.f1c6fe1e214e6p-6,
        0x1.6e8f5826053ccp-6,
        0x1.1c0b647b67e48p-6,
        0x1.cf844e5f3bc52p-7,
        0x1.5058ab6f69f19p-7,
        0x1.fcea21358375cp-7,
        -0x1.c9afa6dfde62cp-8,
        0x1.cac497b043fe3p-6};
    double p = C[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 9; i >= 0; i--) p = fma(p, z, C[i]);
    const double r = fma(s * z, p, s);
    const double a = big ? (x < 0 ? 3.141592653589793 - 2.0 * r : 2.0 * r) : 1.5707963267948966 - r;
    return a * 57.29577951308232;
}
// float degrees of acos(c) exactly as the reference rounds them ((float)(acos((double)c) * 180.0 / M_PI) with the
// library acos), or false when the fast evaluation cannot certify the rounding (then the caller takes the
// reference's own sequence).  The reference's double is within 2^-50 of the true value, acos_deg_fast within
// 2^-46: if the whole 2^-44 interval rounds to one float, that float is the reference's.
FCZ_HD bool acos_deg_certified(float c, float* deg) {
    const double v = acos_deg_fast(c);
    const float lo = (float)(v * (1.0 - 0x1p-44)), hi = (float)(v * (1.0 + 0x1p-44));
    *deg = lo;
    return lo == hi;  // false for NaN
}
// the reference's sequence for one angle, incl. the NaN rule of torsions (torsion_angle.cpp:74-79)
static FCZ_HD_SLOW float angle_deg_slow(float c, bool is_torsion) {
    const double ac = acos((double)c);
    if (is_torsion && ac != ac) return (c < 0) ? 180.0f : 0.0f;
    return (float)(ac * 180.0 / M_PI);
}

// ---- single-precision arccosine in degrees with a proven error bound: the FLOAT-FIRST path of the encoder.
// The exact sequence above is needed only for the few items whose quantised value (or min / max candidacy) a float
// estimate cannot decide; everything else is decided from acosdeg_f and the bound FCZ_ACOS_E0.
//   asin(s) [deg] = s (K + z P(z)),  z = s^2 <= 1/4,  P a degree-4 fit (4.9e-7 deg);  |c| <= 1/2: 90 - asin(c);
//   |c| > 1/2: 2 asin(sqrt((1-|c|)/2)), mirrored for c < 0.  `rs` is an approximation of 1/sqrt((1-|c|)/2) that may be
//   off by a few ulp (MUFU.RSQ on the device, 1/sqrtf on the host); only IEEE operations otherwise, so host and device
//   agree bit for bit given the same rs.
// FCZ_ACOS_E0 bounds |acosdeg_f(c, rs) - (float)(acos((double)c) * 180.0 / M_PI)| over EVERY float c in [-1, 1] and every
// rs within 4 ulp of the true value (exhaustive sweep: tests/test_fastpath.py::test_float_acos_bound, measured 3.82e-5).
#define FCZ_ACOS_E0 8.0e-5f
FCZ_HD float acosdeg_f(float c, float rs) {
    const float ax = fabsf(c);
    const bool big = ax > 0.5f;
    const float zb = (1.0f - ax) * 0.5f;  // exact for 1/2 <= |c| <= 1
    const float z = big ? zb : c * c;
    const float s = big ? zb * rs : ax;
    float p = 0x1.1833680000000p+1f;
    p = fma_(p, z, 0x1.849c500000000p+0f);
    p = fma_(p, z, 0x1.4a1a280000000p+1f);
    p = fma_(p, z, 0x1.12f9e00000000p+2f);
    p = fma_(p, z, 0x1.3193de0000000p+3f);
    const float r = fma_(s * z, p, s * 57.29577951308232f);  // asin(s) in degrees
    if (big) return c < 0.0f ? fma_(-2.0f, r, 180.0f) : 2.0f * r;
    return c < 0.0f ? 90.0f + r : 90.0f - r;
}

// An angle (bond angle, or torsion with its sign) as the float-first path sees it: the value the reference would
// store is within ang_eps(x) of the returned x.
//   * cosine from ONE MUFU.RSQ and a multiply: |c - c_ref| <= 4.5e-7 |c| (2 ulp + two roundings + the reference's own);
//   * where that is not good enough -- |c| > FCZ_COS_HARD, i.e. within 0.81 degrees of 0 / 180, where the reference's
//     angle is a coarse staircase of its float cosine -- the reference's EXACT cosine (cos_ref: certified double
//     sequence) is taken, so the only error left is acosdeg_f's;
//   * |c_ref| > 1: the reference's acos is NaN; torsions then take 0 / 180 (src/torsion_angle.cpp:74-79) -- exact
//     here too; a NaN bond angle, a NaN cosine or a squared-length product outside [1e-30, 1e30] sets `bad`: the chain
//     leaves the float-first path altogether (degenerate input).
#define FCZ_COS_HARD 0.9999f   // acos = 0.81029 degrees
#define FCZ_DEG_HARD 0.81f     // an x closer than this to 0 / 180 was computed from the exact cosine
#define FCZ_ANG_E1 2.5e-3f     // >= 57.2958 * 4.5e-7 * 90 (1 + margin): cosine error over sin(theta) >= d / 90
FCZ_HD float ang_fast(DotParts dp, bool is_tor, bool neg, uint32_t& bad) {
    float c = dp.inner * rsqrt_(dp.p);
    if (!(dp.p >= 1e-30f && dp.p <= 1e30f)) bad |= 1u;
    if (!(fabsf(c) <= FCZ_COS_HARD)) {
        c = cos_ref(dp);
        if (!(fabsf(c) <= 1.0f)) {
            if (!is_tor) bad |= 1u;
            const float t = c < 0.0f ? 180.0f : 0.0f;
            return neg ? -t : t;
        }
    }
    const float zb = (1.0f - fabsf(c)) * 0.5f;
    const float x = acosdeg_f(c, rsqrt_(fmaxf(zb, 1e-30f)));
    return (is_tor && neg) ? -x : x;
}
// bound on |x - reference value| for an x returned by ang_fast (a function of x alone, so that later passes can
// recompute it from the stored value): d = distance to 0 / 180 degrees
FCZ_HD float ang_eps(float x) {
    const float ax = fabsf(x);
    const float d = fminf(ax, 180.0f - ax);
    return d >= FCZ_DEG_HARD ? fma_(FCZ_ANG_E1, rcp_(d), FCZ_ACOS_E0) : FCZ_ACOS_E0;
}
// monotone map float -> int32 (for non-NaN values): integer min / max / atomics on floats
FCZ_HD int32_t ford(float f) {
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    const int32_t i = (int32_t)u;
    return i ^ ((i >> 31) & 0x7fffffff);
}
FCZ_HD float funord(int32_t i) {
    const uint32_t u = (uint32_t)(i ^ ((i >> 31) & 0x7fffffff));
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// reference: angle, src/float3d.h:55-65 -- bond angle at a2 in degrees (double acos).
FCZ_HD float bond_angle_deg(f3 a1, f3 a2, f3 a3) {
    float c = cos_theta(sub3(a1, a2), sub3(a3, a2));
    return (float)(acos((double)c) * 180.0 / M_PI);
}

// reference: getTorsionFromXYZ, src/torsion_angle.cpp:49-94 -- one dihedral in degrees.
FCZ_HD float dihedral_deg(f3 a1, f3 a2, f3 a3, f3 a4) {
    f3 d1 = sub3(a2, a1), d2 = sub3(a3, a2), d3 = sub3(a4, a3);
    f3 u1 = cross3(d1, d2), u2 = cross3(d2, d3);
    float c = cos_theta(u1, u2);
    double ac = acos((double)c);
    float t;
    if (ac != ac) {  // torsion_angle.cpp:74-79: NaN acos -> 180 or 0
        t = (c < 0) ? 180.0f : 0.0f;
    } else {
        t = (float)(ac * 180.0 / M_PI);
    }
    f3 pb = cross3(u2, d2);  // torsion_angle.cpp:87-92
    if ((u1.x * pb.x) + (u1.y * pb.y) + (u1.z * pb.z) < 0) t = -t;
    return t;
}

// ---------------------------------------------------------------- decode-side NeRF (tolerance parity)
// Decoding cannot be bit-identical to the reference (its sincosf comes from glibc), so it is held to
// the BASELINE.md tolerance instead and is free to use fused multiply-adds, rsqrt and a propagated
// frame.  Everything below is exact-arithmetic equivalent to Nerf::place_atom (src/nerf.cpp:39-104).

FCZ_HD float dotf(f3 a, f3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
FCZ_HD f3 crossf(f3 a, f3 b) {
    return mk3(fma_(a.y, b.z, -(b.y * a.z)), fma_(a.z, b.x, -(b.z * a.x)), fma_(a.x, b.y, -(b.x * a.y)));
}
FCZ_HD f3 scalef(f3 v, float s) { return mk3(v.x * s, v.y * s, v.z * s); }
// a + s*v
FCZ_HD f3 axpy(float s, f3 v, f3 a) { return mk3(fma_(s, v.x, a.x), fma_(s, v.y, a.y), fma_(s, v.z, a.z)); }

// degrees -> radians: one double multiply (src/nerf.cpp:63-64 does x*M_PI/180.0 in double)
FCZ_HD float deg2rad(float deg) { return (float)((double)deg * (M_PI / 180.0)); }

// (cos, sin) of an angle
struct cs {
    float c, s;
};
FCZ_HD_SLOW cs cossin_deg_slow(float deg) {
    float r = deg2rad(deg);
    cs o;
#if defined(__CUDA_ARCH__)
    sincosf(r, &o.s, &o.c);
#else
    o.s = sinf(r);
    o.c = cosf(r);
#endif
    return o;
}
// Decode only (the encoder never takes a sine or cosine of an angle).  Quadrant reduction in DEGREES, where it is
// exact: q = rint(deg/90), r = deg - 90q in [-45, 45] without rounding; then the classic single-precision
// minimax polynomials on |x| <= pi/4 (absolute error < 1.2e-7, the same as sincosf gives on a rounded radian
// argument; checked against double in tests/test_fastpath.py).
FCZ_HD cs cossin_deg(float deg) {
    if (!(fabsf(deg) <= 720.0f)) return cossin_deg_slow(deg);  // never for a blob written by an encoder
    const float q = rintf(deg * (1.0f / 90.0f));
    const float r = fma_(q, -90.0f, deg);
    const float x = r * 0.017453292519943295f;
    const float z = x * x;
    const float sp = fma_(fma_(fma_(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * x, x);
    const float cp = fma_(fma_(fma_(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z, fma_(-0.5f, z, 1.0f));
    const int k = (int)q & 3;
    cs o;
    o.s = (k & 1) ? cp : sp;
    o.c = (k & 1) ? sp : cp;
    if (k & 2) o.s = -o.s;
    if ((k + 1) & 2) o.c = -o.c;
    return o;
}
FCZ_HD cs cossin_angle(f3 a, f3 b, f3 c) {
    f3 u = sub3(a, b), v = sub3(c, b);
    cs o;
    o.c = dotf(u, v) * rsqrt_(dotf(u, u) * dotf(v, v));
    float s2 = fma_(-o.c, o.c, 1.0f);
    o.s = s2 > 0.0f ? sqrtf(s2) : 0.0f;
    return o;
}

// The local frame Nerf::place_atom builds from three atoms (a, b, c): bcn along b->c, n normal to
// the a-b-c plane, nbc = n x bcn (src/nerf.cpp:52-85).
struct NerfFrame {
    f3 bcn, n, nbc;
};
FCZ_HD NerfFrame frame_from(f3 a, f3 b, f3 c) {
    NerfFrame f;
    f3 ab = sub3(b, a), bc = sub3(c, b);
    f.bcn = scalef(bc, rsqrt_(dotf(bc, bc)));
    f3 n = crossf(ab, f.bcn);
    f.n = scalef(n, rsqrt_(dotf(n, n)));
    f.nbc = crossf(f.n, f.bcn);
    return f;
}
// Place the next atom of a chain from the frame of the previous three and advance the frame:
//   D      = c + len * (-cos(th) bcn + cos(tau) sin(th) nbc + sin(tau) sin(th) n)     (nerf.cpp:66-98)
//   bcn'   = (D - c) / len
//   n'     = normalise((c - b) x bcn') =             - sin(tau) nbc + cos(tau) n
//   nbc'   = n' x bcn'                 = -sin(th) bcn - cos(tau) cos(th) nbc - sin(tau) cos(th) n
// i.e. the frame place_atom would rebuild from (b, c, D), obtained by ROTATING the old frame with an
// orthogonal 3x3 matrix: no square root, no division, and -- unlike recomputing nbc' as a cross
// product, which multiplies the vectors' norm errors and blows up exponentially -- rounding errors
// only add up linearly along the chain.
FCZ_HD f3 nerf_step(NerfFrame& f, f3 c, float len, cs ang, cs tor) {
    const float x = -ang.c, y = tor.c * ang.s, z = tor.s * ang.s;
    const float u = -ang.s, v = -(tor.c * ang.c), w = -(tor.s * ang.c);
    const f3 b0 = f.bcn, b1 = f.nbc, b2 = f.n;
    f.bcn = mk3(fma_(z, b2.x, fma_(y, b1.x, x * b0.x)), fma_(z, b2.y, fma_(y, b1.y, x * b0.y)), fma_(z, b2.z, fma_(y, b1.z, x * b0.z)));
    f.nbc = mk3(fma_(w, b2.x, fma_(v, b1.x, u * b0.x)), fma_(w, b2.y, fma_(v, b1.y, u * b0.y)), fma_(w, b2.z, fma_(v, b1.z, u * b0.z)));
    f.n = mk3(fma_(tor.c, b2.x, -(tor.s * b1.x)), fma_(tor.c, b2.y, -(tor.s * b1.y)), fma_(tor.c, b2.z, -(tor.s * b1.z)));
    return axpy(len, f.bcn, c);
}
// The same step for ONE Cartesian component: the update combines the frame vectors with scalar coefficients,
// so x, y and z evolve independently and three lanes can share a chain (bit-identical to nerf_step).
struct NerfFrame1 {
    float bcn, nbc, n;
};
FCZ_HD float comp3(f3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
FCZ_HD float nerf_step1(NerfFrame1& f, float c, float len, cs ang, cs tor) {
    const float x = -ang.c, y = tor.c * ang.s, z = tor.s * ang.s;
    const float u = -ang.s, v = -(tor.c * ang.c), w = -(tor.s * ang.c);
    const float b0 = f.bcn, b1 = f.nbc, b2 = f.n;
    f.bcn = fma_(z, b2, fma_(y, b1, x * b0));
    f.nbc = fma_(w, b2, fma_(v, b1, u * b0));
    f.n = fma_(tor.c, b2, -(tor.s * b1));
    return fma_(len, f.bcn, c);
}
// One-off placement from three arbitrary predecessor atoms (side chains, src/nerf.cpp:106-155)
FCZ_HD f3 place_from(f3 a, f3 b, f3 c, float len, cs ang, cs tor) {
    NerfFrame f = frame_from(a, b, c);
    return nerf_step(f, c, len, ang, tor);
}

// ------------------------------------------------------------------ discretiser (src/discretizer.cpp)

// double -> unsigned as x86-64 gcc emits it (cvttsd2si to int64, low 32 bits); NaN/out of range -> 0.
FCZ_HD unsigned d2u(double x) {
    if (!(x > -9.2e18 && x < 9.2e18)) return 0u;
    return (unsigned)(long long)x;
}
// src/discretizer.cpp:28-32 factors from (min,max); n_bin converted to float
FCZ_HD float disc_factor(float mn, float mx, unsigned nb) { return (float)nb / (mx - mn); }
FCZ_HD float cont_factor(float mn, float mx, unsigned nb) { return (mx - mn) / (float)nb; }
// src/discretizer.cpp:49 rounding variant (+0.5 in double)
FCZ_HD unsigned disc_round(float x, float mn, float disc_f) { return d2u((double)((x - mn) * disc_f) + 0.5); }
// src/discretizer.cpp:55-57 truncating scalar variant (side chains, src/foldcomp.cpp:532-538)
FCZ_HD unsigned disc_trunc(float x, float mn, float disc_f) { return d2u((double)((x - mn) * disc_f)); }
// src/discretizer.cpp:64,71 / src/foldcomp.cpp:155-158: two float roundings
FCZ_HD float continuize(unsigned q, float mn, float cont_f) { return ((float)q * cont_f) + mn; }

// FixedAngleDiscretizer(255), src/discretizer.h:89-106: min=-180, max=180
FCZ_HD float sc_min() { return (float)-180.0; }
FCZ_HD float sc_disc_f() { return (float)255u / ((float)180.0 - (float)-180.0); }
FCZ_HD float sc_cont_f() { return ((float)180.0 - (float)-180.0) / (float)255u; }

// std::min_element / max_element combine step with their NaN behaviour (src/discretizer.cpp:27-28):
// a NaN never replaces the running value, and a NaN FIRST element is never replaced.  `first` is
// handled by the caller (result = NaN when v[0] is NaN); here NaNs are simply ignored.
FCZ_HD float min_ignore_nan(float a, float b) { return (b < a) ? b : a; }
FCZ_HD float max_ignore_nan(float a, float b) { return (a < b) ? b : a; }

}  // namespace fcz
#endif  // FCZ_MATH_H
