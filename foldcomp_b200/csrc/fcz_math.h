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
#else
#define FCZ_HD inline
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

// degrees -> radians exactly as src/nerf.cpp:63-64 (double multiply, double divide, to float)
FCZ_HD float deg2rad(float deg) { return (float)((double)deg * M_PI / 180.0); }

// (cos, sin) of an angle given in degrees, the pair place_atom needs
struct cs {
    float c, s;
};
FCZ_HD cs cossin_deg(float deg) {
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

// reference: Nerf::place_atom, src/nerf.cpp:39-104, with the trigonometry hoisted out:
// ang = (cos, sin) of the bond angle, tor = (cos, sin) of the torsion.  Operation order of the
// remaining float arithmetic is the reference's.
FCZ_HD f3 place_atom(f3 a, f3 b, f3 c, float len, cs ang, cs tor) {
    f3 ab = sub3(b, a), bc = sub3(c, b);
    float bc_norm = norm3(bc);
    f3 bcn = mk3(bc.x / bc_norm, bc.y / bc_norm, bc.z / bc_norm);
    float d2x = (-1 * len) * ang.c;
    float d2y = (len * tor.c) * ang.s;
    float d2z = (len * tor.s) * ang.s;
    f3 n = cross3(ab, bcn);
    float n_norm = norm3(n);
    n = mk3(n.x / n_norm, n.y / n_norm, n.z / n_norm);
    f3 nbc = cross3(n, bcn);
    f3 d;
    d.x = ((bcn.x * d2x + nbc.x * d2y) + n.x * d2z) + c.x;
    d.y = ((bcn.y * d2x + nbc.y * d2y) + n.y * d2z) + c.y;
    d.z = ((bcn.z * d2x + nbc.z * d2y) + n.z * d2z) + c.z;
    return d;
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
