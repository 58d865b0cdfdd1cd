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

FCZ_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
FCZ_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
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
