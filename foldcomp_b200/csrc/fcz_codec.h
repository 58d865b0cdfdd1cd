// foldcomp_b200/csrc/fcz_codec.h -- per-chain FCZ encode / decode written against an abstract
// execution context (host + device).
//
// `encode_chain` and `decode_chain` are the whole hot path for ONE chain, expressed as phases of
// strided loops separated by block barriers.  The CUDA kernels (fcz_engine.cu) instantiate them
// with a CTA context (256 threads, shared-memory workspace, TMA-staged inputs); the CPU tests
// instantiate them with a one-thread context (tests/emu/) to check the algorithm against the
// oracle without a GPU.  All pointers are generic: the caller decides whether a workspace array
// or the blob/coordinates live in shared or global memory.
//
// Reference call stacks replaced:
//   encode: Foldcomp::preprocess src/foldcomp.cpp:450-559, compress 562-606, writeStream 1038-1109
//   decode: Foldcomp::read src/foldcomp.cpp:904-1036, decompress 779-902,
//           reconstructBackboneAtoms 167-246, reconstructBackboneReverse 248-273,
//           Nerf::reconstructWithReversed src/nerf.cpp:342-379, weightedAverage
//           src/atom_coordinate.cpp:145-163, Nerf::reconstructAminoAcid src/nerf.cpp:106-155
#ifndef FCZ_CODEC_H
#define FCZ_CODEC_H

#include <stdlib.h>
#include <string.h>

#include "../../include/fcz_engine.h"
#include "fcz_format.h"

namespace fcz {

// Device-friendly copy of the residue tables (built once from fcz_tables.h).  FCZ_CODE_ROWS = 32 rows: a record's
// residue field has five bits, and the reference maps every value its switch does not know (24..31) to UNK
// (convertIntToOneLetterCode / convertIntToThreeLetterCode default branch, src/utility.cpp:297-377, 461-), so rows
// 24..31 repeat the UNK row: a raw code from an untrusted blob indexes these tables safely and decodes as the
// reference decodes it.  Encode INPUT codes must be < FCZ_NUM_CODES (the plans check).
// (FCZ_CODE_ROWS and norm_code live in fcz_format.h.)
struct Tables {
    // (the encode kernel keeps a copy of the fields up to `alt` in shared memory: natoms, name1, pred)
    uint8_t natoms[FCZ_CODE_ROWS];
    uint8_t name1[FCZ_CODE_ROWS];  // one-letter codes (header firstResidue / lastResidue)
    uint16_t pred[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];
    uint8_t alt[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];
    // side-chain byte as a step function of cos(torsion) (see build_tables), both tables increasing:
    // non-negated torsion: byte = 127 + #{i: -c >= sc_pos[i]};  negated torsion: byte = #{i: c >= sc_neg[i]}
    float sc_pos[128];
    float sc_neg[128];
    float blen[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];
    cs bang[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];  // (cos, sin) of the table bond angle
    cs sc_tor[256];  // (cos, sin) of every side-chain torsion byte: FixedAngleDiscretizer(255).continuize(b)
    // one 16-byte entry per (residue code, slot) for the decoder's side-chain placement: predecessors, bond length and
    // (cos, sin) of the bond angle come with ONE 128-bit load instead of three look-ups
    struct alignas(16) ScEntry { uint32_t pred; float blen; cs bang; } sc_ent[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];
};

// The side-chain byte the reference stores for a torsion whose cosine is the float c
// (src/torsion_angle.cpp:70-92, src/discretizer.cpp:55-57 via src/foldcomp.cpp:532-538), host libm.
inline unsigned sc_byte_of_cos(float c, bool neg) {
    double ac = acos((double)c);
    float t = (ac != ac) ? ((c < 0) ? 180.0f : 0.0f) : (float)(ac * 180.0 / M_PI);
    if (neg) t = -t;
    return disc_trunc(t, sc_min(), sc_disc_f()) & 0xFFu;
}
inline float float_from_ordered(int32_t k) {  // monotone bijection int32 <-> float (excluding NaN)
    uint32_t u = k >= 0 ? (uint32_t)k : ~((uint32_t)k) ^ 0x80000000u;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline int32_t ordered_from_float(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? (int32_t)~(u ^ 0x80000000u) : (int32_t)u;
}

// Filled on the host with host libm: the table angles are compile-time constants of the format,
// so their (cos,sin) are computed once, exactly as cossin_deg would.
inline void build_tables(Tables* t) {
    // Side-chain thresholds in cos-space.  The byte is a monotone step function of the float c (acos is
    // strictly monotone at float-input granularity, every later step is monotone), so each step's boundary
    // is found by bisection over the floats in [-2, 2] (|c| > 1 takes the reference's NaN branch).
    {
        const int64_t lo0 = ordered_from_float(-2.0f), hi0 = ordered_from_float(2.0f);
        for (int i = 0; i < 128; i++) {
            // non-negated torsion: byte >= 128+i holds for small c; find the LARGEST such c
            const unsigned j = 128u + (unsigned)i;
            if (sc_byte_of_cos(-2.0f, false) < j) t->sc_pos[i] = 3.0f;
            else if (sc_byte_of_cos(2.0f, false) >= j) t->sc_pos[i] = -3.0f;
            else {
                int64_t lo = lo0, hi = hi0;  // predicate true at lo, false at hi
                while (hi - lo > 1) {
                    int64_t mid = lo + (hi - lo) / 2;
                    if (sc_byte_of_cos(float_from_ordered((int32_t)mid), false) >= j) lo = mid; else hi = mid;
                }
                t->sc_pos[i] = -float_from_ordered((int32_t)lo);
            }
            // negated torsion: byte >= 1+i holds for large c; find the SMALLEST such c
            const unsigned jn = 1u + (unsigned)i;
            if (i == 127 || sc_byte_of_cos(2.0f, true) < jn) t->sc_neg[i] = 3.0f;
            else if (sc_byte_of_cos(-2.0f, true) >= jn) t->sc_neg[i] = -3.0f;
            else {
                int64_t lo = lo0, hi = hi0;  // predicate false at lo, true at hi
                while (hi - lo > 1) {
                    int64_t mid = lo + (hi - lo) / 2;
                    if (sc_byte_of_cos(float_from_ordered((int32_t)mid), true) >= jn) hi = mid; else lo = mid;
                }
                t->sc_neg[i] = float_from_ordered((int32_t)hi);
            }
        }
    }
    for (int row = 0; row < FCZ_CODE_ROWS; row++) {
        const int c = (int)norm_code((unsigned)row);
        t->natoms[row] = FCZ_NATOMS[c];
        if (natoms_packed((unsigned)row) != (uint32_t)FCZ_NATOMS[c]) abort();  // the packed constants of fcz_format.h went stale
        t->name1[row] = (uint8_t)FCZ_NAME1[c];
        for (int k = 0; k < FCZ_MAX_ATOMS; k++) {
            t->alt[row][k] = FCZ_ALT[c][k];
            t->pred[row][k] = FCZ_PRED[c][k];
            t->blen[row][k] = FCZ_BLEN[c][k];
            float r = (float)((double)FCZ_BANG[c][k] * M_PI / 180.0);
            t->bang[row][k].c = cosf(r);
            t->bang[row][k].s = sinf(r);
            t->sc_ent[row][k].pred = FCZ_PRED[c][k];
            t->sc_ent[row][k].blen = t->blen[row][k];
            t->sc_ent[row][k].bang = t->bang[row][k];
        }
    }
    for (int b = 0; b < 256; b++) {  // src/foldcomp.cpp:338-369 + src/nerf.cpp:64,66-70
        float r = deg2rad(continuize((unsigned)b, sc_min(), sc_cont_f()));
        t->sc_tor[b].c = cosf(r);
        t->sc_tor[b].s = sinf(r);
    }
}

// ------------------------------------------------------------------------------------------ encode

// scratch floats needed by the min/max reduction: per-warp partials + results + discretiser params
#define FCZ_RED_FLOATS(nwarps) ((nwarps) * 14 + 14 + 21)

// Scratch words of the float-first backbone path (EncChain::fl):
//   [FL_N]      number of list entries appended (may exceed the capacity: then the chain takes the exact path)
//   [FL_BAD]    != 0: degenerate geometry seen, the chain takes the exact path
//   [FL_BND+4a .. +3]  array a (header order), as ordered ints: min and max over the items of (x - eps), min and max of (x + eps)
//   [FL_EX+2a, +1]     exact min / max over the re-evaluated items of array a (a = 6: the B-factors)
// The list of undecided values (EncChain::list, value index a * L + r) and their exact values (EncChain::xe) hold
// list_cap entries each: about 3 % of a chain's 6 (L-1) values end up there, enc_list_cap leaves a factor ~1.5.
// Both may alias EncChain::sres, which is dead once phase 2 is over (the list is first touched after the barrier
// that ends phase 3).
enum { FL_N = 0, FL_BAD = 1, FL_BND = 2, FL_EX = 26, FCZ_FL_WORDS = 40 };
FCZ_HD uint32_t enc_list_cap(uint32_t max_res) { return max_res <= 1024u ? 256u : (max_res + 3u) / 4u; }

struct EncChain {
    uint32_t L, A, title_len;
    int32_t b;               // anchor threshold
    const uint8_t* type;     // [L] residue codes
    const float* bfac;       // [L]
    const float* X;          // [3A] atoms of this chain (staged copy or global)
    const char* title;       // [title_len]
    const fcz_chain_meta* meta;
    uint8_t* B;              // blob destination (staged copy or global), Layout.size bytes
    // workspace
    uint32_t* aoff;          // [L+1] first atom of each residue, relative to the chain
    uint16_t* sres;          // [A-3L] residue of each side-chain atom (= side-chain torsion)
    float* ang;              // [6*L] the six backbone arrays, header order, stride L: angles, then (in place) their quantised values
    float* red;              // [FCZ_RED_FLOATS(nwarps)]
    uint32_t* fl;            // [FCZ_FL_WORDS]
    const Tables* tbg;       // the complete tables (global memory): threshold tables of the exact side-chain path
    uint32_t* list;          // [list_cap] undecided values of the float-first path
    float* xe;               // [list_cap] their exact values
    uint32_t list_cap;
};

// The side-chain byte for an EXACT cosine c (the reference's float): count of thresholds in cos-space, found from a
// cheap estimate corrected against the table (exact whatever the estimate's quality).
FCZ_HD uint8_t sc_byte_exact(const Tables* tb, float c, bool neg) {
    const float* tz = neg ? tb->sc_neg : tb->sc_pos;
    if (c != c) c = 2.0f;  // NaN cosine: the reference's acos is NaN and c < 0 is false -> 0 degrees
    c = c > 2.0f ? 2.0f : (c < -2.0f ? -2.0f : c);  // |c| > 1 all take the NaN branch: the table spans [-2, 2]
    const float z = neg ? c : -c;
    // estimate (Abramowitz & Stegun 4.4.45): acos(x) ~ sqrt(1-x) (a0 + a1 x + a2 x^2 + a3 x^3), 0 <= x <= 1
    const float x = fminf(fabsf(c), 1.0f);
    float a = sqrtf(1.0f - x) * (1.5707288f + x * (-0.2121144f + x * (0.0742610f + x * -0.0187293f))) * 57.29578f;
    if (c < 0.0f) a = 180.0f - a;
    int n = (int)(((neg ? -a : a) + 180.0f) * sc_disc_f()) - (neg ? 0 : 127);
    n = n < 0 ? 0 : (n > 128 ? 128 : n);
    while (n > 0 && !(z >= tz[n - 1])) n--;   // exact count of thresholds <= z (tables increase)
    while (n < 128 && z >= tz[n]) n++;
    return (uint8_t)(neg ? n : 127 + n);
}
static FCZ_HD_SLOW uint8_t sc_byte_slow(const Tables* tb, float inner, float p, bool neg) {
    return sc_byte_exact(tb, cos_exact_slow(inner, p), neg);
}
// Side-chain byte from the parts of cos(torsion) and the torsion's sign (see encode_chain phase 2): the reference stores
// (unsigned)((t - (-180.f)) * (255 / 360.f)) -- float add, float multiply, truncation (src/discretizer.cpp:55-57 via
// src/foldcomp.cpp:532-538).  Both are monotone, so the float-first estimate x of t with |x - t| <= eps brackets the
// product; when both ends truncate to the same integer that is the reference's byte, otherwise (an estimate within
// eps of a byte boundary: ~2e-4 of the items) the reference's exact cosine goes through the threshold table.
FCZ_HD uint8_t sc_byte_fast(const Tables* tb, DotParts dp, bool neg) {
    uint32_t bad = 0;
    const float x = ang_fast(dp, true, neg, bad);
    const float e = ang_eps(x);
    const float t_lo = ((x - e) + 180.0f) * sc_disc_f(), t_hi = ((x + e) + 180.0f) * sc_disc_f();
    const int n_lo = (int)t_lo, n_hi = (int)t_hi;
    if (!bad && n_lo == n_hi) return (uint8_t)n_lo;
    return sc_byte_slow(tb, dp.inner, dp.p, neg);
}

FCZ_HD f3 bb_atom(const EncChain& ch, uint32_t j) {  // j-th backbone atom (N,CA,C = slots 0..2)
    uint32_t r = j / 3u, k = j - 3u * r;
    return ld3(ch.X + 3u * (ch.aoff[r] + k));
}

// The NaN an x86-64 build of the reference leaves in a bond angle at backbone atom m (angle(), src/float3d.h:55-65, over
// getCosineTheta 36-43 and the library acos) -- needed only when that angle is the FIRST of its array, because then it
// becomes the header's min / cont_f and reaches the file (src/discretizer.cpp:27-32).  SSE arithmetic returns a NaN
// operand quieted (payload and sign kept); an invalid operation on non-NaN operands (0/0 for coincident atoms,
// inf - inf, inf/inf for infinite coordinates) returns the default NaN 0xFFC00000; conversions float <-> double and
// glibc's acos ((x-x)/(x-x)) keep both.  So: the quieted first NaN among the nine input floats, else the default NaN.
// (With several DIFFERENT NaN payloads among the nine inputs the winner depends on the compiler's operand order --
// two builds of the reference disagree with each other there; the scan order below is one valid choice.)
FCZ_HD float x86_angle_nan(const EncChain& ch, uint32_t m) {
    const uint32_t order[3] = {m - 1u, m, m + 1u};
    for (int i = 0; i < 3; i++) {
        const f3 p = bb_atom(ch, order[i]);
        if (p.x != p.x) return u2f(f2u(p.x) | 0x00400000u);
        if (p.y != p.y) return u2f(f2u(p.y) | 0x00400000u);
        if (p.z != p.z) return u2f(f2u(p.z) | 0x00400000u);
    }
    return u2f(0xFFC00000u);
}

// Discretiser parameters (min, disc_f, cont_f) of one array from its extremes and its first element
// (Discretizer::Discretizer, src/discretizer.cpp:22-33), k = array index (6 = B-factors).
FCZ_HD void enc_params(const EncChain& ch, int k, float lo, float hi, float first, float* prm) {
    if (first != first) {  // min_element/max_element keep a NaN first element
        // A bond angle that came out NaN: WHICH NaN is hardware business (see x86_angle_nan); B-factors are input
        // and are copied bit for bit (this->min = *std::min_element(...) is a plain move).
        if (k >= 3 && k < 6) first = x86_angle_nan(ch, k == A_CACN ? 2u : (k == A_CNCA ? 3u : 4u));
        lo = first; hi = first;
    }
    const unsigned nb = (k < 6) ? n_bins(k) : 255u;
    prm[0] = lo;
    prm[1] = disc_factor(lo, hi, nb);
    prm[2] = cont_factor(lo, hi, nb);
    if (prm[2] != prm[2]) {
        // (max - min) is NaN: x86 propagates the quieted NaN operand (min = max = the NaN first element) through
        // the subtraction and the division, and produces its default NaN (negative, 0xFFC00000) for inf - inf;
        // a GPU would put its canonical 0x7FFFFFFF into the header instead (src/discretizer.cpp:27-32).
        const float nanv = u2f(lo != lo ? (f2u(lo) | 0x00400000u) : 0xFFC00000u);
        prm[1] = nanv;
        prm[2] = nanv;
    }
}

// One backbone value by the reference's exact sequence: array a (header order) at index r -- the torsion over backbone
// atoms 3r+k .. 3r+k+3 or the bond angle at backbone atom 3r+k+2, k = 0 (psi, CA-C-N), 1 (omega, C-N-CA), 2 (phi, N-CA-C)
// (src/torsion_angle.cpp:49-94, src/nerf.cpp:495-508, split src/foldcomp.cpp:488-505).
FCZ_HD void bb_item_atoms(const EncChain& ch, uint32_t r, uint32_t k, f3& p0, f3& p1, f3& p2, f3& p3) {
    const uint32_t a0 = ch.aoff[r], a1 = ch.aoff[r + 1u];
    // backbone atoms j .. j+3 = (r,k) (r,k+1) ... wrapping into residue r+1
    const uint32_t i0 = a0 + k, i1 = (k < 2u) ? a0 + k + 1u : a1, i2 = (k < 1u) ? a0 + 2u : a1 + (k - 1u), i3 = a1 + k;
    p0 = ld3(ch.X + 3u * i0); p1 = ld3(ch.X + 3u * i1); p2 = ld3(ch.X + 3u * i2); p3 = ld3(ch.X + 3u * i3);
}
FCZ_HD uint32_t bb_array_k(uint32_t a) { return (a == A_PSI || a == A_CACN) ? 0u : ((a == A_OMEGA || a == A_CNCA) ? 1u : 2u); }
// torsion over p0..p3 (is_tor) or bond angle at p2 over p1, p2, p3, by the reference's exact sequence
static FCZ_HD_SLOW float bb_value_of_atoms(bool is_tor, f3 p0, f3 p1, f3 p2, f3 p3) {
    const f3 d2 = sub3(p2, p1), d3 = sub3(p3, p2);
    f3 v1, v2;
    bool neg = false;
    if (is_tor) {
        const f3 d1 = sub3(p1, p0);
        v1 = cross3(d1, d2);
        v2 = cross3(d2, d3);
        const f3 pb = cross3(v2, d2);
        neg = (v1.x * pb.x) + (v1.y * pb.y) + (v1.z * pb.z) < 0;
    } else {
        v1 = sub3(p1, p2);
        v2 = d3;
    }
    const float c = cos_ref(dot_parts(v1, v2));
    float deg;
    if (!acos_deg_certified(c, &deg)) deg = angle_deg_slow(c, is_tor);  // ~2e-6 of items, and |c| > 1
    return (is_tor && neg) ? -deg : deg;
}
static FCZ_HD_SLOW float bb_value_exact(const EncChain& ch, uint32_t a, uint32_t r) {
    f3 p0, p1, p2, p3;
    bb_item_atoms(ch, r, bb_array_k(a), p0, p1, p2, p3);
    return bb_value_of_atoms(a < 3u, p0, p1, p2, p3);
}

// Backbone angles of one chain BEFORE quantisation, six floats per residue r: the torsions over backbone atoms 3r+j ..
// 3r+j+3 (psi_r, omega_r, phi_r+1; zero for the last residue) and the bond angles at backbone atoms 3r, 3r+1, 3r+2 (zero
// at the chain's first and last atom) -- Foldcomp::preprocess's backboneTorsionAngles (getTorsionFromXYZ,
// src/torsion_angle.cpp:46-96) and backboneBondAngles (Nerf::getBondAngles, src/nerf.cpp:495-508), src/foldcomp.cpp:484-496;
// what the CPython get_data(pdb_text) returns (foldcomp/foldcomp.cxx:633-671).  ch needs L, X and aoff.
template <class Ctx>
FCZ_HD void enc_raw_angles(Ctx& cx, const EncChain& ch, float* out) {
    const uint32_t L = ch.L;
    for (uint32_t e = (uint32_t)cx.tid; e < 6u * L; e += (uint32_t)cx.nthr) {
        const uint32_t r = e / 6u, j = e - 6u * r;
        float v = 0.0f;
        if (j < 3u) {
            if (r + 1u < L) v = bb_value_exact(ch, j == 0u ? A_PSI : (j == 1u ? A_OMEGA : A_PHI), r);
        } else {
            const uint32_t m = 3u * r + (j - 3u);  // backbone atom the angle sits at
            if (m == 1u) {
                const uint32_t a0 = ch.aoff[0];
                const f3 z = {0.f, 0.f, 0.f};
                v = bb_value_of_atoms(false, z, ld3(ch.X + 3u * a0), ld3(ch.X + 3u * (a0 + 1u)), ld3(ch.X + 3u * (a0 + 2u)));
            } else if (m >= 2u && m + 1u < 3u * L) {
                const uint32_t k = (m - 2u) % 3u;
                v = bb_value_exact(ch, k == 0u ? A_CACN : (k == 1u ? A_CNCA : A_NCAC), (m - 2u) / 3u);
            }
        }
        out[e] = v;
    }
}

// q = floor(t + 0.5) as the reference's (unsigned)((double)t + 0.5) gives it for 0 <= t < 2^23, in exact float steps
FCZ_HD uint32_t round_half_up(float t) {
    const int n = (int)t;
    return (uint32_t)n + ((t - (float)n) >= 0.5f ? 1u : 0u);
}

// ---- the backbone arrays by the reference's exact sequence for EVERY item (the fallback of the float-first path:
// degenerate geometry, or more undecided items than the list holds): phase 3 computes all 6(L-1) values, phase 4
// reduces them, then they are quantised in place.
template <class Ctx>
FCZ_HD void enc_backbone_exact(Ctx& cx, const EncChain& ch) {
    const uint32_t L = ch.L;
    {
        const uint32_t nV = 6u * (L - 1u);
        for (uint32_t w = cx.tid; w < nV; w += cx.nthr) {
            const uint32_t a = w / (L - 1u), r = w - a * (L - 1u);
            ch.ang[a * L + r] = bb_value_exact(ch, a, r);
        }
    }
    cx.sync();
    {
        float mn[6], mx[6];
        for (int k = 0; k < 6; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
        for (uint32_t i = cx.tid; i + 1u < L; i += cx.nthr) {
            for (int k = 0; k < 6; k++) {
                float v = ch.ang[k * L + i];
                mn[k] = min_ignore_nan(mn[k], v);
                mx[k] = max_ignore_nan(mx[k], v);
            }
        }
        for (int k = 0; k < 6; k++) {
            mn[k] = cx.wmin(mn[k]);
            mx[k] = cx.wmax(mx[k]);
        }
        if (cx.lane == 0) {
            for (int k = 0; k < 6; k++) {
                ch.red[cx.warp * 14 + k] = mn[k];
                ch.red[cx.warp * 14 + 7 + k] = mx[k];
            }
        }
        cx.sync();
        float* res = ch.red + cx.nwarps * 14;  // [14] unused, then [21] (min, disc_f, cont_f) x 7
        for (int k = cx.tid; k < 6; k += cx.nthr) {
            float lo = INFINITY, hi = -INFINITY;
            for (int w = 0; w < cx.nwarps; w++) {
                lo = min_ignore_nan(lo, ch.red[w * 14 + k]);
                hi = max_ignore_nan(hi, ch.red[w * 14 + 7 + k]);
            }
            enc_params(ch, k, lo, hi, ch.ang[k * L], res + 14 + 3 * k);
        }
        cx.sync();
    }
    const float* prm = ch.red + cx.nwarps * 14 + 14;
    uint32_t* q = reinterpret_cast<uint32_t*>(ch.ang);
    for (uint32_t w = cx.tid; w < 6u * (L - 1u); w += cx.nthr) {
        const uint32_t a = w / (L - 1u), r = w - a * (L - 1u);
        q[a * L + r] = disc_round(ch.ang[a * L + r], prm[3 * a], prm[3 * a + 1]);
    }
}

template <class Ctx>
FCZ_HD void encode_chain(Ctx& cx, const Tables* tb, const EncChain& ch) {
    const uint32_t L = ch.L, A = ch.A;
    const int n_anchor = anchor_count(L, ch.b);
    const Layout y = make_layout(L, A - 3u * L, ch.title_len, (uint32_t)n_anchor);
    uint8_t* B = ch.B;
    uint32_t* fl = ch.fl;
    int32_t* fli = reinterpret_cast<int32_t*>(ch.fl);
    float* flf = reinterpret_cast<float*>(ch.fl);

    // ---- phase 1: residue -> first atom (exclusive scan of table atom counts), side-chain atom -> residue
    {
        const uint32_t chunk = (L + cx.nthr - 1) / cx.nthr;
        uint32_t r0 = cx.tid * chunk; if (r0 > L) r0 = L;
        uint32_t r1 = r0 + chunk; if (r1 > L) r1 = L;
        uint32_t sum = 0;
        for (uint32_t r = r0; r < r1; r++) sum += natoms_packed(ch.type[r]);
        uint32_t base = cx.excl_scan(sum);
        for (uint32_t r = r0; r < r1; r++) {
            ch.aoff[r] = base;
            uint32_t n = natoms_packed(ch.type[r]);
            for (uint32_t k = 3u; k < n; k++) ch.sres[base - 3u * r + (k - 3u)] = (uint16_t)r;
            base += n;
        }
        if (r1 == L) ch.aoff[L] = base;
        // scratch of the float-first path: counters, bounds (ordered ints), exact extremes
        for (uint32_t i = cx.tid; i < (uint32_t)FCZ_FL_WORDS; i += cx.nthr) {
            uint32_t v = 0u;
            if (i >= (uint32_t)FL_BND && i < (uint32_t)FL_EX) {
                const uint32_t w = (i - FL_BND) & 3u;  // min(x-eps), max(x-eps), min(x+eps), max(x+eps)
                v = (uint32_t)ford((w & 1u) ? -INFINITY : INFINITY);
            } else if (i >= (uint32_t)FL_EX) {
                v = (uint32_t)ford(((i - FL_EX) & 1u) ? -INFINITY : INFINITY);
            }
            fl[i] = v;
        }
    }
    cx.stage_wait();  // coordinates staged by the caller are now visible
    cx.sync();
    cx.mark(0);  // E_SCAN

    // ---- phase 2: side-chain bytes, one dihedral per side-chain atom (src/sidechain.cpp:149-168 +
    // FixedAngleDiscretizer, src/foldcomp.cpp:532-538), float first (sc_byte_fast).
    {
        const uint32_t S = A - 3u * L;
        for (uint32_t t = cx.tid; t < S; t += cx.nthr) {
            const uint32_t r = ch.sres[t];
            const uint32_t a0 = ch.aoff[r];
            const uint32_t k = 3u + (t - (a0 - 3u * r));
            const uint32_t w = a0 + k;
            const unsigned pr = tb->pred[ch.type[r]][k];
            const f3 p0 = ld3(ch.X + 3u * (a0 + (pr & 15u))), p1 = ld3(ch.X + 3u * (a0 + ((pr >> 4) & 15u)));
            const f3 p2 = ld3(ch.X + 3u * (a0 + ((pr >> 8) & 15u))), p3 = ld3(ch.X + 3u * w);
            const f3 d1 = sub3(p1, p0), d2 = sub3(p2, p1), d3 = sub3(p3, p2);
            const f3 u1 = cross3(d1, d2), u2 = cross3(d2, d3);
            const f3 pb = cross3(u2, d2);  // torsion_angle.cpp:87-92
            const bool neg = (u1.x * pb.x) + (u1.y * pb.y) + (u1.z * pb.z) < 0;
            B[y.o_sc + t] = sc_byte_fast(ch.tbg, dot_parts(u1, u2), neg);
        }
    }
    cx.mark(1);  // E_SIDE (no barrier: timing only)

#if !defined(FCZ_ENC_BACKBONE_PAIR)  // (FCZ_ENC_BACKBONE_PAIR: experiment build with one item per residue pair, libfcz_engine_pair.so)
    // ---- phase 3: backbone, float first.  One item = residue pair (r, r+1) and k in {0,1,2}: the torsion over the four
    // backbone atoms 3r+k .. 3r+k+3 (psi, omega, phi of record r: src/torsion_angle.cpp:49-94, src/foldcomp.cpp:488-492)
    // AND the bond angle at the third of them (CA-C-N, C-N-CA, N-CA-C of record r: src/nerf.cpp:495-508,
    // src/foldcomp.cpp:496-505), which share their bond vectors.  Every value is stored as a float estimate x (ang_fast);
    // per array the extremes of x - eps and x + eps are gathered (thread, warp, then integer atomics on ordered floats).
    // k is uniform over a warp when the warps come in threes.
    {
        const int kstep = (cx.nwarps % 3 == 0) ? 3 : 1;
        const uint32_t rstart = (uint32_t)cx.lane + (uint32_t)cx.wsize * (uint32_t)(kstep == 3 ? cx.warp / 3 : cx.warp);
        const uint32_t rstep = (uint32_t)cx.wsize * (uint32_t)(kstep == 3 ? cx.nwarps / 3 : cx.nwarps);
        uint32_t bad = 0;
        for (uint32_t k = (kstep == 3) ? (uint32_t)(cx.warp % 3) : 0u; k < 3u; k += (uint32_t)kstep) {
            const uint32_t at = (k == 0u) ? (uint32_t)A_PSI : (k == 1u ? (uint32_t)A_OMEGA : (uint32_t)A_PHI);
            const uint32_t ab = (k == 0u) ? (uint32_t)A_CACN : (k == 1u ? (uint32_t)A_CNCA : (uint32_t)A_NCAC);
            float tlo_mn = INFINITY, tlo_mx = -INFINITY, thi_mn = INFINITY, thi_mx = -INFINITY;
            float blo_mn = INFINITY, blo_mx = -INFINITY, bhi_mn = INFINITY, bhi_mx = -INFINITY;
            for (uint32_t r = rstart; r + 1u < L; r += rstep) {
                f3 p0, p1, p2, p3;
                bb_item_atoms(ch, r, k, p0, p1, p2, p3);
                const f3 d1 = sub3(p1, p0), d2 = sub3(p2, p1), d3 = sub3(p3, p2);
                const f3 u1 = cross3(d1, d2), u2 = cross3(d2, d3);
                const f3 pb = cross3(u2, d2);
                const bool neg = (u1.x * pb.x) + (u1.y * pb.y) + (u1.z * pb.z) < 0;
                const float xt = ang_fast(dot_parts(u1, u2), true, neg, bad);
                // bond angle at p2 between p1 - p2 = -d2 and p3 - p2 = d3: negation is exact, so the reference's
                // products and sums (src/float3d.h:36-43) are these with the sign flipped
                DotParts da;
                da.inner = -((d2.x * d3.x) + (d2.y * d3.y) + (d2.z * d3.z));
                da.p = (d2.x * d2.x + d2.y * d2.y + d2.z * d2.z) * (d3.x * d3.x + d3.y * d3.y + d3.z * d3.z);
                const float xb = ang_fast(da, false, false, bad);
                ch.ang[at * L + r] = xt;
                ch.ang[ab * L + r] = xb;
                const float et = ang_eps(xt), eb = ang_eps(xb);
                const float tl = xt - et, th = xt + et, bl = xb - eb, bh = xb + eb;
                tlo_mn = fminf(tlo_mn, tl); tlo_mx = fmaxf(tlo_mx, tl); thi_mn = fminf(thi_mn, th); thi_mx = fmaxf(thi_mx, th);
                blo_mn = fminf(blo_mn, bl); blo_mx = fmaxf(blo_mx, bl); bhi_mn = fminf(bhi_mn, bh); bhi_mx = fmaxf(bhi_mx, bh);
            }
            int32_t v[8] = {ford(tlo_mn), ford(tlo_mx), ford(thi_mn), ford(thi_mx), ford(blo_mn), ford(blo_mx), ford(bhi_mn), ford(bhi_mx)};
            for (int i = 0; i < 8; i++) v[i] = (i & 1) ? cx.wmax_i(v[i]) : cx.wmin_i(v[i]);
            if (cx.lane == 0) {
                for (int i = 0; i < 8; i++) {
                    int32_t* dst = fli + FL_BND + 4 * (i < 4 ? at : ab) + (i & 3);
                    if (i & 1) cx.atomic_max_i(dst, v[i]); else cx.atomic_min_i(dst, v[i]);
                }
            }
        }
#else
    // ---- phase 3: backbone, float first.  One item = one residue pair (r, r+1): the three torsions over its six
    // backbone atoms (psi, omega, phi of record r: src/torsion_angle.cpp:49-94, src/foldcomp.cpp:488-492) and the three bond
    // angles at C, N', CA' (CA-C-N, C-N-CA, N-CA-C of record r: src/nerf.cpp:495-508, src/foldcomp.cpp:496-505) share five
    // bond vectors and four cross products.  Every value is stored as a float estimate x (ang_fast); per array the
    // extremes of x - eps and x + eps are gathered (warp reduction on ordered integers, then shared-memory atomics).
    {
        uint32_t bad = 0;
        for (uint32_t r0 = 0; r0 + 1u < L; r0 += (uint32_t)cx.nthr) {
            const uint32_t r = r0 + (uint32_t)cx.tid;
            const bool on = r + 1u < L;
            float x[6];  // header order: phi psi omega N-CA-C CA-C-N C-N-CA
            if (on) {
                const uint32_t a0 = ch.aoff[r], a1 = ch.aoff[r + 1u];
                const f3 n0 = ld3(ch.X + 3u * a0), ca0 = ld3(ch.X + 3u * a0 + 3u), c0 = ld3(ch.X + 3u * a0 + 6u);
                const f3 n1 = ld3(ch.X + 3u * a1), ca1 = ld3(ch.X + 3u * a1 + 3u), c1 = ld3(ch.X + 3u * a1 + 6u);
                const f3 d0 = sub3(ca0, n0), d1 = sub3(c0, ca0), d2 = sub3(n1, c0), d3 = sub3(ca1, n1), d4 = sub3(c1, ca1);
                const f3 c01 = cross3(d0, d1), c12 = cross3(d1, d2), c23 = cross3(d2, d3), c34 = cross3(d3, d4);
                const float s01 = c01.x * c01.x + c01.y * c01.y + c01.z * c01.z, s12 = c12.x * c12.x + c12.y * c12.y + c12.z * c12.z;
                const float s23 = c23.x * c23.x + c23.y * c23.y + c23.z * c23.z, s34 = c34.x * c34.x + c34.y * c34.y + c34.z * c34.z;
                DotParts dp;
                {   // psi: atoms N CA C N'
                    const f3 pb = cross3(c12, d1);
                    const bool neg = (c01.x * pb.x) + (c01.y * pb.y) + (c01.z * pb.z) < 0;
                    dp.inner = (c01.x * c12.x) + (c01.y * c12.y) + (c01.z * c12.z); dp.p = s01 * s12;
                    x[A_PSI] = ang_fast(dp, true, neg, bad);
                }
                {   // omega: atoms CA C N' CA'
                    const f3 pb = cross3(c23, d2);
                    const bool neg = (c12.x * pb.x) + (c12.y * pb.y) + (c12.z * pb.z) < 0;
                    dp.inner = (c12.x * c23.x) + (c12.y * c23.y) + (c12.z * c23.z); dp.p = s12 * s23;
                    x[A_OMEGA] = ang_fast(dp, true, neg, bad);
                }
                {   // phi: atoms C N' CA' C'
                    const f3 pb = cross3(c34, d3);
                    const bool neg = (c23.x * pb.x) + (c23.y * pb.y) + (c23.z * pb.z) < 0;
                    dp.inner = (c23.x * c34.x) + (c23.y * c34.y) + (c23.z * c34.z); dp.p = s23 * s34;
                    x[A_PHI] = ang_fast(dp, true, neg, bad);
                }
                // bond angle at atom b between a - b and c - b, i.e. between -d and d': negation is exact, so the
                // reference's products and sums (src/float3d.h:36-43) are these with the sign flipped
                const float q1 = d1.x * d1.x + d1.y * d1.y + d1.z * d1.z, q2 = d2.x * d2.x + d2.y * d2.y + d2.z * d2.z;
                const float q3 = d3.x * d3.x + d3.y * d3.y + d3.z * d3.z, q4 = d4.x * d4.x + d4.y * d4.y + d4.z * d4.z;
                dp.inner = -((d1.x * d2.x) + (d1.y * d2.y) + (d1.z * d2.z)); dp.p = q1 * q2;
                x[A_CACN] = ang_fast(dp, false, false, bad);
                dp.inner = -((d2.x * d3.x) + (d2.y * d3.y) + (d2.z * d3.z)); dp.p = q2 * q3;
                x[A_CNCA] = ang_fast(dp, false, false, bad);
                dp.inner = -((d3.x * d4.x) + (d3.y * d4.y) + (d3.z * d4.z)); dp.p = q3 * q4;
                x[A_NCAC] = ang_fast(dp, false, false, bad);
                for (int a = 0; a < 6; a++) ch.ang[a * L + r] = x[a];
            }
            for (int a = 0; a < 6; a++) {
                float lo = INFINITY, hi = -INFINITY;  // an idle lane changes no extreme
                float lo2 = -INFINITY, hi2 = INFINITY;
                if (on) {
                    const float e = ang_eps(x[a]);
                    lo = x[a] - e; hi = x[a] + e; lo2 = lo; hi2 = hi;
                }
                const int32_t v0 = cx.wmin_i(ford(lo)), v1 = cx.wmax_i(ford(lo2)), v2 = cx.wmin_i(ford(hi2)), v3 = cx.wmax_i(ford(hi));
                if (cx.lane == 0) {
                    int32_t* dst = fli + FL_BND + 4 * a;
                    cx.atomic_min_i(dst, v0); cx.atomic_max_i(dst + 1, v1); cx.atomic_min_i(dst + 2, v2); cx.atomic_max_i(dst + 3, v3);
                }
            }
        }
#endif
        if (bad) fl[FL_BAD] = 1u;
        // B-factors: exact extremes (they are input)
        {
            float mn = INFINITY, mx = -INFINITY;
            for (uint32_t i = cx.tid; i < L; i += cx.nthr) {
                const float v = ch.bfac[i];
                mn = min_ignore_nan(mn, v);
                mx = max_ignore_nan(mx, v);
            }
            const int32_t imn = cx.wmin_i(ford(mn)), imx = cx.wmax_i(ford(mx));
            if (cx.lane == 0) {
                cx.atomic_min_i(fli + FL_EX + 12, imn);
                cx.atomic_max_i(fli + FL_EX + 13, imx);
            }
        }
    }
    cx.sync();
    cx.mark(2);  // E_BACKBONE

    float* prm = ch.red + cx.nwarps * 14 + 14;  // [21] (min, disc_f, cont_f) x 7
    if (cx.tid == 6 % cx.nthr) enc_params(ch, 6, funord(fli[FL_EX + 12]), funord(fli[FL_EX + 13]), ch.bfac[0], prm + 18);
    bool exact_path = fl[FL_BAD] != 0u;
    if (!exact_path) {
        // ---- phase 4a: decide every value from its estimate.  The true min of an array lies in [mL, mU] = [min(x-eps),
        // min(x+eps)], its max in [ML, MU]; disc_f = RN(nb / RN(max - min)) is monotone in both, so it lies between
        // f_lo <= RN(nb / RN(MU - mL)) and f_hi >= RN(nb / RN(ML - mU)) -- taken from ONE MUFU.RCP each, widened by 1e-6
        // (rcp_ is good to 2 ulp, the product to half an ulp; every thread derives them itself: no extra pass).
        // x in [x-eps, x+eps]: RN(x - min) * disc_f is bracketed by t_lo = RN(RN(lo - mU) * f_lo) and
        // t_hi = RN(RN(hi - mL) * f_hi) (monotone IEEE operations on non-negative values).  Same rounded integer at both
        // ends and no chance of being the array's min or max: that integer is the reference's; otherwise the value
        // goes on the list.  (Crossed bounds give a negative or infinite f_hi: nothing is decided, the list overflows.)
        {
            uint32_t* q = reinterpret_cast<uint32_t*>(ch.ang);
            const uint32_t n1 = L - 1u;
            for (uint32_t a = 0; a < 6u; a++) {
                const float mL = funord(fli[FL_BND + 4 * a]), ML = funord(fli[FL_BND + 4 * a + 1]);
                const float mU = funord(fli[FL_BND + 4 * a + 2]), MU = funord(fli[FL_BND + 4 * a + 3]);
                const float nb = (float)n_bins((int)a);
                const float f_lo = (nb * rcp_(MU - mL)) * (1.0f - 1e-6f), f_hi = (nb * rcp_(ML - mU)) * (1.0f + 1e-6f);
                for (uint32_t r = cx.tid; r < n1; r += cx.nthr) {
                    const float x = ch.ang[a * L + r];
                    const float e = ang_eps(x);
                    const float lo = x - e, hi = x + e;
                    const float t_lo = (lo - mU) * f_lo, t_hi = (hi - mL) * f_hi;
                    uint32_t q_lo = 0u;
                    bool decided = false;
                    if (lo > mU && hi < ML && t_lo >= 0.0f && t_hi < 65536.0f) {
                        q_lo = round_half_up(t_lo);
                        decided = q_lo == round_half_up(t_hi);
                    }
                    if (decided) {
                        q[a * L + r] = q_lo;
                    } else {
                        const uint32_t j = cx.atomic_add(&fl[FL_N], 1u);
                        if (j < ch.list_cap) ch.list[j] = a * L + r;
                    }
                }
            }
        }
        cx.sync();
        cx.mark(5);  // E_DECIDE
        exact_path = fl[FL_N] > ch.list_cap;
    }
    if (!exact_path) {
        // ---- phase 4b: the undecided values by the reference's exact sequence; their extremes are the arrays' extremes
        const uint32_t n = fl[FL_N];
        for (uint32_t j = cx.tid; j < n; j += cx.nthr) {
            const uint32_t v = ch.list[j], a = v / L, r = v - a * L;
            const float x = bb_value_exact(ch, a, r);
            ch.xe[j] = x;
            if (x == x) {
                cx.atomic_min_i(fli + FL_EX + 2 * a, ford(x));
                cx.atomic_max_i(fli + FL_EX + 2 * a + 1, ford(x));
            } else {
                fl[FL_BAD] = 1u;  // cannot happen (a NaN bond angle sets `bad` in phase 3); belt and braces
            }
        }
        cx.sync();
        cx.mark(6);  // E_LIST
        exact_path = fl[FL_BAD] != 0u;
    }
    if (!exact_path) {
        // ---- phase 4c: the discretiser parameters from the exact extremes (no NaN on this path: the first element plays no
        // role), by every thread that needs them -- the list's threads for their value, thread 0 for the header
        const uint32_t n = fl[FL_N];
        uint32_t* q = reinterpret_cast<uint32_t*>(ch.ang);
        for (uint32_t j = cx.tid; j < n; j += cx.nthr) {
            const uint32_t v = ch.list[j], a = v / L;
            const float lo = funord(fli[FL_EX + 2 * a]), hi = funord(fli[FL_EX + 2 * a + 1]);
            float pa[3];
            enc_params(ch, (int)a, lo, hi, lo, pa);
            q[v] = disc_round(ch.xe[j], pa[0], pa[1]);
        }
        if (cx.tid == 0) {
            for (int a = 0; a < 6; a++) {
                const float lo = funord(fli[FL_EX + 2 * a]), hi = funord(fli[FL_EX + 2 * a + 1]);
                enc_params(ch, a, lo, hi, lo, prm + 3 * a);
            }
        }
    } else {
        cx.sync();  // (phase-4 scratch reads above are done)
        enc_backbone_exact(cx, ch);
    }
    cx.sync();
    cx.mark(7);  // E_QUANT

    // ---- phase 5: serialise (src/foldcomp.cpp:1038-1109)
    if (cx.tid == 0) {
        B[0] = 'F'; B[1] = 'C'; B[2] = 'M'; B[3] = 'P';
        put_u16(B + OFF_NRES, L);
        put_u16(B + OFF_NATOM, ch.meta->n_atom);
        put_u16(B + OFF_IDXRES, ch.meta->idx_residue);
        put_u16(B + OFF_IDXATOM, ch.meta->idx_atom);
        B[OFF_NANCHOR] = (uint8_t)n_anchor;
        B[OFF_CHAIN] = ch.meta->chain;
        B[14] = 0; B[15] = 0;  // struct padding: uninitialised in the reference, zero here
        put_u32(B + OFF_NSC, y.n_sc);
        B[OFF_FIRSTRES] = tb->name1[ch.type[0]];      // src/foldcomp.cpp:467
        B[OFF_LASTRES] = tb->name1[ch.type[L - 1u]];  // src/foldcomp.cpp:468
        B[22] = 0; B[23] = 0;
        put_u32(B + OFF_LENTITLE, ch.title_len);
        for (int k = 0; k < 6; k++) {
            put_f32(B + OFF_MINS + 4 * k, prm[3 * k]);
            put_f32(B + OFF_CONTFS + 4 * k, prm[3 * k + 2]);
        }
        uint8_t* o = B + y.o_oxt;  // src/foldcomp.cpp:1061-1064
        o[0] = ch.meta->has_oxt;
        put_f32(o + 1, ch.meta->oxt[0]);
        put_f32(o + 5, ch.meta->oxt[1]);
        put_f32(o + 9, ch.meta->oxt[2]);
        put_f32(B + y.o_temp, prm[18]);      // tempFactorsDisc.min     (src/foldcomp.cpp:1098-1100)
        put_f32(B + y.o_temp + 4, prm[20]);  // tempFactorsDisc.cont_f
    }
    for (uint32_t i = cx.tid; i < (uint32_t)n_anchor; i += cx.nthr)
        put_u32(B + y.o_aidx + 4u * i, (uint32_t)anchor_index(L, n_anchor, (int)i));
    for (uint32_t i = cx.tid; i < ch.title_len; i += cx.nthr) B[y.o_title + i] = (uint8_t)ch.title[i];
    // anchor atoms: N, CA, C of each anchor residue = 9 consecutive floats (src/foldcomp.cpp:1051-1059)
    for (uint32_t e = cx.tid; e < 9u * (uint32_t)n_anchor; e += cx.nthr) {
        uint32_t i = e / 9u, w = e - 9u * i;
        uint32_t r = (uint32_t)anchor_index(L, n_anchor, (int)i);
        put_f32(B + y.o_anchor + 4u * e, ch.X[3u * ch.aoff[r] + w]);
    }
    // backbone records (src/foldcomp.cpp:581-602) and B-factor bytes (src/foldcomp.cpp:1102-1107)
    {
        const uint32_t* qv = reinterpret_cast<const uint32_t*>(ch.ang);
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
            unsigned q[6] = {0, 0, 0, 0, 0, 0};
            if (r + 1u < L) {
                for (int k = 0; k < 6; k++) q[k] = qv[k * L + r];
            }
            pack_record(B + y.o_rec + 8u * r, ch.type[r], q[A_PHI], q[A_PSI], q[A_OMEGA], q[A_NCAC], q[A_CACN], q[A_CNCA]);
            B[y.o_temp + 8u + r] = (uint8_t)disc_round(ch.bfac[r], prm[18], prm[19]);
        }
    }
    cx.sync();
    cx.mark(3);  // E_PACK
}

// ------------------------------------------------------------------------------------------ decode

// floats of per-segment scratch: S[9] true start atoms, T[12] rigid transform local->true, TAIL[9] local
// coordinates of the segment's last residue, F[12] frame after the first placed residue in local
// coordinates (bcn, nbc, n) + its origin (the C atom), HEAD[6] local N',CA' of the first placed residue,
// RF[12] frame (bcn, nbc, n) and last atom of the reverse pass once it has placed atom 3
// A[9] the stored anchor that opens the segment (as aligned floats), I[3] first / last residue index of
// the segment (uint32 bits) and 1/(atoms in the segment).  Slot n_seg holds only A (the closing anchor).
// CS[13] (cos,sin) of the bond angles and torsions of the segment's first record + its N-CA length (so that
// the stitch needs nothing but this scratch).
#define FCZ_SEG_FLOATS 88
enum { SEG_S = 0, SEG_T = 9, SEG_TAIL = 21, SEG_F = 30, SEG_HEAD = 42, SEG_RF = 48, SEG_A = 60, SEG_I = 69, SEG_CS = 72 };

struct DecChain {
    const uint8_t* blob;  // staged copy or global
    Layout y;
    int use_alt;
    // outputs
    float* out_xyz;       // [3A] canonical slot order; staged copy or global.  Also the backbone work area.
    uint8_t* out_type;    // [L]
    float* out_bfac;      // [L]
    fcz_chain_meta* out_meta;
    char* out_title;      // [title_len] or NULL
    // workspace
    uint32_t* aoff;       // [L+1]
    cs* tor;              // [3(L-1)] (cos,sin) of psi,omega,phi per record
    cs* ang;              // [3(L-1)] (cos,sin) of CA-C-N, C-N-CA, N-CA-C per record
    float* seg;           // [n_anchor * FCZ_SEG_FLOATS]
    float* rev;           // [9L] reverse-pass backbone atoms (true coordinates)
    const float* loc;     // [9L] forward-pass (local) backbone atoms when they do not live in out_xyz, else NULL
    uint8_t* segid;       // [L] anchor segment that owns (emits) each residue
    uint16_t* order;      // [L] residues sorted by atom count (side chains), with bins[32] u32 of scratch; or NULL
    uint32_t* bins;
    const uint8_t* codes;  // [L] residue codes when a copy is at hand (else read from the records), or NULL
    const uint8_t* sc;     // [A - 3L] side-chain torsion bytes when a copy is at hand (else read from the blob), or NULL
};

// y = R x + t, T = rows of R (9) then t (3)
FCZ_HD f3 xform(const float* T, f3 x) {
    return mk3(fma_(T[2], x.z, fma_(T[1], x.y, fma_(T[0], x.x, T[9]))), fma_(T[5], x.z, fma_(T[4], x.y, fma_(T[3], x.x, T[10]))),
               fma_(T[8], x.z, fma_(T[7], x.y, fma_(T[6], x.x, T[11]))));
}
FCZ_HD f3 get_f3(const uint8_t* p) { return mk3(get_f32(p), get_f32(p + 4), get_f32(p + 8)); }
FCZ_HD uint32_t seg_a0(const float* sg) { return f2u(sg[SEG_I]); }
FCZ_HD uint32_t seg_a1(const float* sg) { return f2u(sg[SEG_I + 1]); }
FCZ_HD float n_ca_len(unsigned code) { return code == FCZ_CODE_PRO ? FCZ_PRO_N_TO_CA : FCZ_N_TO_CA; }

// weightedAverage (src/atom_coordinate.cpp:145-163): atom i of n: (fwd*(n-i) + rev*i) / n
FCZ_HD f3 blend(f3 fwd, f3 rev, float wf, float wr, float inv_n) {
    return mk3(fma_(rev.x, wr, fwd.x * wf) * inv_n, fma_(rev.y, wr, fwd.y * wf) * inv_n, fma_(rev.z, wr, fwd.z * wf) * inv_n);
}

// The decode of one chain is five phases.  decode_chain() runs them back to back inside one execution context (the
// CPU model of tests/emu/ does); the engine (fcz_engine.cu) runs them as kernels over all chains of a length tier --
// unpack + passes in k_dec_front, the stitch thread-per-chain in k_dec_stitch_t, blend + side chains in k_dec_back --
// so that the serial stitch costs its latency once per batch instead of once per CTA.

template <class Ctx>
FCZ_HD void dec_unpack(Ctx& cx, const Tables* tb, const DecChain& ch) {
    const Layout& y = ch.y;
    const uint32_t L = y.L;
    const uint8_t* blob = ch.blob;
    const uint8_t* rec = blob + y.o_rec;
    const int n_seg = (int)y.n_anchor - 1;
    const uint32_t nT = 3u * L - 3u;
    (void)L; (void)rec; (void)n_seg; (void)nT; (void)blob; (void)tb;
    // ---- phase 1: records -> residue codes, atom offsets (exclusive scan), (cos,sin) of the continuised angles
    // (convertBytesToBackboneChain src/foldcomp.cpp:60-77, decompressBackboneChain 122-153,
    // _continuize 155-158; the deg->rad and sincos of Nerf::place_atom src/nerf.cpp:63-70 are hoisted
    // here so the recurrences below carry no transcendental)
    {
        float mins[6], cfs[6];
        for (int k = 0; k < 6; k++) {
            mins[k] = get_f32(blob + OFF_MINS + 4 * k);
            cfs[k] = get_f32(blob + OFF_CONTFS + 4 * k);
        }
        const uint32_t chunk = (L + cx.nthr - 1) / cx.nthr;
        uint32_t r0 = cx.tid * chunk; if (r0 > L) r0 = L;
        uint32_t r1 = r0 + chunk; if (r1 > L) r1 = L;
        uint32_t sum = 0;
        for (uint32_t r = r0; r < r1; r++) {
            sum += natoms_packed(rec[8u * r] >> 3);
        }
        uint32_t base = cx.excl_scan(sum);
        for (uint32_t r = r0; r < r1; r++) {
            ch.aoff[r] = base;
            base += natoms_packed(rec[8u * r] >> 3);
        }
        if (r1 == L) ch.aoff[L] = base;
        const float tmin = get_f32(blob + y.o_temp), tcf = get_f32(blob + y.o_temp + 4);
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
            Record q = unpack_record(rec + 8u * r);
            ch.out_type[r] = (uint8_t)norm_code(q.res);  // 24..31 decode as UNK (see Tables)
            ch.out_bfac[r] = continuize(blob[y.o_temp + 8u + r], tmin, tcf);  // src/foldcomp.cpp:884-886
            if (r + 1u < L) {
                ch.tor[3u * r + 0u] = cossin_deg(continuize(q.psi, mins[A_PSI], cfs[A_PSI]));
                ch.tor[3u * r + 1u] = cossin_deg(continuize(q.omg, mins[A_OMEGA], cfs[A_OMEGA]));
                ch.tor[3u * r + 2u] = cossin_deg(continuize(q.phi, mins[A_PHI], cfs[A_PHI]));
                ch.ang[3u * r + 0u] = cossin_deg(continuize(q.cac, mins[A_CACN], cfs[A_CACN]));
                ch.ang[3u * r + 1u] = cossin_deg(continuize(q.cnc, mins[A_CNCA], cfs[A_CNCA]));
                ch.ang[3u * r + 2u] = cossin_deg(continuize(q.nca, mins[A_NCAC], cfs[A_NCAC]));
            }
        }
        // anchors as aligned floats and the segment bounds, so that the serial phases never touch the
        // unaligned blob bytes (anchor atoms: src/foldcomp.cpp:925-953)
        for (uint32_t e = cx.tid; e < 9u * y.n_anchor; e += cx.nthr) {
            const uint32_t i = e / 9u;
            ch.seg[i * FCZ_SEG_FLOATS + SEG_A + (e - 9u * i)] = get_f32(blob + y.o_anchor + 4u * e);
        }
        for (uint32_t i = cx.tid; i + 1u < y.n_anchor; i += cx.nthr) {
            const uint32_t a0 = get_u32(blob + y.o_aidx + 4u * i), a1 = get_u32(blob + y.o_aidx + 4u * (i + 1u));
            float* sg = ch.seg + i * FCZ_SEG_FLOATS;
            sg[SEG_I] = u2f(a0);
            sg[SEG_I + 1] = u2f(a1);
            sg[SEG_I + 2] = 1.0f / (float)(3u * (a1 - a0 + 1u));
        }
        if (cx.tid == 0) {
            fcz_chain_meta m;
            m.n_atom = (uint16_t)get_u16(blob + OFF_NATOM);
            m.idx_residue = (uint16_t)get_u16(blob + OFF_IDXRES);
            m.idx_atom = (uint16_t)get_u16(blob + OFF_IDXATOM);
            m.chain = blob[OFF_CHAIN];
            m.has_oxt = blob[y.o_oxt];
            m.oxt[0] = get_f32(blob + y.o_oxt + 1);
            m.oxt[1] = get_f32(blob + y.o_oxt + 5);
            m.oxt[2] = get_f32(blob + y.o_oxt + 9);
            *ch.out_meta = m;
        }
        if (ch.out_title)
            for (uint32_t i = cx.tid; i < y.title_len; i += cx.nthr) ch.out_title[i] = (char)blob[y.o_title + i];
    }
}

// One work item of phase 2: component k of the forward (dir 0) or reverse (dir 1) pass of segment s; see dec_passes.
FCZ_HD void dec_pass_item(const Tables* tb, const DecChain& ch, int dir, int s, int k) {
    (void)tb;
    const uint8_t* rec = ch.blob + ch.y.o_rec;
    float* sg = ch.seg + s * FCZ_SEG_FLOATS;
    const uint32_t a0 = seg_a0(sg), a1 = seg_a1(sg);
    if (dir == 0) {
        const f3 p0 = ld3(sg + SEG_A), p1 = ld3(sg + SEG_A + 3), p2 = ld3(sg + SEG_A + 6);
        const NerfFrame F = frame_from(p0, p1, p2);
        NerfFrame1 f = {comp3(F.bcn, k), comp3(F.nbc, k), comp3(F.n, k)};
        float q0 = comp3(p0, k), q1 = comp3(p1, k), q2 = comp3(p2, k);
        for (uint32_t r = a0; r < a1; r++) {
            const uint32_t t = 3u * r;
            q0 = nerf_step1(f, q2, FCZ_C_TO_N, ch.ang[t], ch.tor[t]);                             // N
            q1 = nerf_step1(f, q0, n_ca_len(rec[8u * r] >> 3), ch.ang[t + 1u], ch.tor[t + 1u]);  // CA
            q2 = nerf_step1(f, q1, FCZ_CA_TO_C, ch.ang[t + 2u], ch.tor[t + 2u]);                  // C
            float* o = ch.out_xyz + 3u * ch.aoff[r + 1u] + k;
            o[0] = q0; o[3] = q1; o[6] = q2;
            if (k == 0) ch.segid[r] = (uint8_t)s;
            if (r == a0) {  // frame carried by the first placed residue, origin at its C
                sg[SEG_F + k] = f.bcn; sg[SEG_F + 3 + k] = f.nbc; sg[SEG_F + 6 + k] = f.n; sg[SEG_F + 9 + k] = q2;
                sg[SEG_HEAD + k] = q0; sg[SEG_HEAD + 3 + k] = q1;
                if (k == 0) {
                    for (int j = 0; j < 3; j++) {  // what the stitch needs of the first record
                        sg[SEG_CS + 2 * j] = ch.ang[t + j].c; sg[SEG_CS + 2 * j + 1] = ch.ang[t + j].s;
                        sg[SEG_CS + 6 + 2 * j] = ch.tor[t + j].c; sg[SEG_CS + 6 + 2 * j + 1] = ch.tor[t + j].s;
                    }
                    sg[SEG_CS + 12] = n_ca_len(rec[8u * r] >> 3);
                }
            }
        }
        sg[SEG_TAIL + k] = q0; sg[SEG_TAIL + 3 + k] = q1; sg[SEG_TAIL + 6 + k] = q2;
    } else if (a1 > a0) {
        const float* anc = sg + FCZ_SEG_FLOATS + SEG_A;
        // reversed chain starts as the stored anchor: a = C, b = CA, c = N
        const f3 pn = ld3(anc);
        const NerfFrame F = frame_from(ld3(anc + 6), ld3(anc + 3), pn);
        NerfFrame1 f = {comp3(F.bcn, k), comp3(F.nbc, k), comp3(F.n, k)};
        float rc = comp3(pn, k);
        // atoms q = n-4 .. 3 are C, CA, N of residues a1-1 .. a0+1; g = backbone atom index in the chain
        for (uint32_t r = a1 - 1u; r > a0; r--) {
            const uint32_t g = 3u * r;
            const cs b2 = ch.ang[g + 1u], b1 = ch.ang[g], b0 = ch.ang[g - 1u];  // angle at atom g+k+1
            const cs t2 = ch.tor[g + 2u], t1 = ch.tor[g + 1u], t0 = ch.tor[g];  // torsion g+k
            const float c = nerf_step1(f, rc, FCZ_C_TO_N, b2, t2);
            const float ca = nerf_step1(f, c, FCZ_CA_TO_C, b1, t1);
            rc = nerf_step1(f, ca, FCZ_N_TO_CA, b0, t0);
            float* o = ch.rev + 3u * g + k;
            o[0] = rc; o[3] = ca; o[6] = c;
        }
        sg[SEG_RF + k] = f.bcn; sg[SEG_RF + 3 + k] = f.nbc; sg[SEG_RF + 6 + k] = f.n; sg[SEG_RF + 9 + k] = rc;
    }
}

template <class Ctx>
FCZ_HD void dec_passes(Ctx& cx, const Tables* tb, const DecChain& ch) {
    const Layout& y = ch.y;
    const uint32_t L = y.L;
    const uint8_t* blob = ch.blob;
    const uint8_t* rec = blob + y.o_rec;
    const int n_seg = (int)y.n_anchor - 1;
    const uint32_t nT = 3u * L - 3u;
    (void)L; (void)rec; (void)n_seg; (void)nT; (void)blob; (void)tb;
    // ---- phase 2: both NeRF passes of every anchor segment, TWO lanes per segment.
    //  even lane: FORWARD pass (reconstructBackboneAtoms, src/foldcomp.cpp:167-246; Pro N-CA length taken
    //   from the record being consumed, 204-212) in the segment's LOCAL frame: it starts from the STORED
    //   anchor instead of the blended tail of the previous segment (not known yet).  From its 4th placed
    //   atom on a forward pass is a rigid body hanging off its first placed residue (N',CA',C'), so the
    //   true pass is this one moved by a rigid transform that phase 3 determines.
    //  odd lane: REVERSE pass (reconstructBackboneReverse src/foldcomp.cpp:248-273, Nerf::reconstructWithReversed
    //   src/nerf.cpp:342-379 with forward indices) from the stored anchor s+1, in true coordinates, for atoms
    //   n-4 .. 3.  The reference feeds it the bond angles re-measured on the forward atoms (getBondAngles,
    //   src/nerf.cpp:495-508); for atoms 4.. those are the stored angles the forward pass was built with
    //   (up to float noise), so the stored values are used and the pass does not wait for the forward one.
    //   Bond lengths by atom kind, never the Pro length (src/nerf.h:37-43).  The last three reverse atoms
    //   depend on the true start atoms and are finished in phase 4.
    // (forward and reverse lanes sit in DIFFERENT warps: lanes of one warp would serialise the two loops)
    // Each pass of a segment is shared by THREE lanes, one per Cartesian component (nerf_step1): work items
    // (direction, segment, component); the first half of the warps takes the forward items, the rest the reverse.
    const int items = 3 * n_seg;
    if (cx.nwarps >= 2) {
        const int half = cx.nwarps / 2;
        const int dir = cx.warp >= half ? 1 : 0;
        const int w = dir ? cx.warp - half : cx.warp, nw = dir ? cx.nwarps - half : half;
        for (int i = w * cx.wsize + cx.lane; i < items; i += nw * cx.wsize) dec_pass_item(tb, ch, dir, i / 3, i - 3 * (i / 3));
    } else {
        for (int dir = 0; dir < 2; dir++)
            for (int i = cx.tid; i < items; i += cx.nthr) dec_pass_item(tb, ch, dir, i / 3, i - 3 * (i / 3));
    }
}

// ---- phase 3: stitch.  Serial over segments (the only cross-segment dependency of the reference,
// src/foldcomp.cpp:855-857: the blended tail of segment s seeds segment s+1).  Per segment: place N',CA',C'
// from the true start atoms, derive the rigid transform local->true from the two frames, move the local tail,
// blend it with the stored anchor (weightedAverage, src/atom_coordinate.cpp:145-163, last three atoms only).
// Works on the segment scratch alone; element (slot s, field f) sits at seg[(s*FCZ_SEG_FLOATS + f) * stride]:
// stride 1 = one chain's scratch, stride C = structure-of-arrays over the C chains of a sub-batch (one thread
// per chain, coalesced across the warp).
// M maps the fields: SegFull = the scratch as the other phases see it; SegPacked = only what the stitch reads, with
// its outputs S, T overwriting TAIL, F of the same slot (dead once the step has loaded them).
struct SegFull { enum { S = SEG_S, T = SEG_T, TAIL = SEG_TAIL, F = SEG_F, A = SEG_A, I = SEG_I, CS = SEG_CS, N = FCZ_SEG_FLOATS }; };
struct SegPacked { enum { I = 0, F = 3, T = 3, TAIL = 15, S = 15, A = 24, CS = 33, N = 46 }; };
template <class M>
FCZ_HD void dec_stitch_core(float* seg, size_t stride, int n_seg) {
#define SEG_S M::S
#define SEG_T M::T
#define SEG_TAIL M::TAIL
#define SEG_F M::F
#define SEG_A M::A
#define SEG_I M::I
#define SEG_CS M::CS
#define SG(s_, f_) seg[((size_t)(s_) * M::N + (size_t)(f_)) * stride]
#define LD3(s_, f_) mk3(SG(s_, f_), SG(s_, (f_) + 1), SG(s_, (f_) + 2))
#define ST3(s_, f_, v_) do { const f3 v__ = (v_); SG(s_, f_) = v__.x; SG(s_, (f_) + 1) = v__.y; SG(s_, (f_) + 2) = v__.z; } while (0)
    f3 s0 = LD3(0, SEG_A), s1 = LD3(0, SEG_A + 3), s2 = LD3(0, SEG_A + 6);
    for (int s = 0; s < n_seg; s++) {
        // all inputs of this step first (independent of the serial chain), then the chain itself
        const uint32_t a0 = f2u(SG(s, SEG_I)), a1 = f2u(SG(s, SEG_I + 1));
        const float inv = SG(s, SEG_I + 2);
        const f3 e0 = LD3(s + 1, SEG_A), e1 = LD3(s + 1, SEG_A + 3), e2 = LD3(s + 1, SEG_A + 6);
        const f3 l1 = LD3(s, SEG_F), l2 = LD3(s, SEG_F + 3), l3 = LD3(s, SEG_F + 6), lo = LD3(s, SEG_F + 9);
        const f3 q0 = LD3(s, SEG_TAIL), q1 = LD3(s, SEG_TAIL + 3), q2 = LD3(s, SEG_TAIL + 6);
        cs b[3], w[3];
        for (int j = 0; j < 3; j++) {
            b[j].c = SG(s, SEG_CS + 2 * j); b[j].s = SG(s, SEG_CS + 2 * j + 1);
            w[j].c = SG(s, SEG_CS + 6 + 2 * j); w[j].s = SG(s, SEG_CS + 6 + 2 * j + 1);
        }
        const float l_nca = SG(s, SEG_CS + 12);
        ST3(s, SEG_S, s0); ST3(s, SEG_S + 3, s1); ST3(s, SEG_S + 6, s2);
        f3 t0 = s0, t1 = s1, t2 = s2;  // forward tail in true coordinates
        if (a1 > a0) {
            NerfFrame fa = frame_from(s0, s1, s2);
            f3 n = nerf_step(fa, s2, FCZ_C_TO_N, b[0], w[0]);
            f3 ca = nerf_step(fa, n, l_nca, b[1], w[1]);
            f3 c = nerf_step(fa, ca, FCZ_CA_TO_C, b[2], w[2]);
            float T[12];
            // R = bcn_a bcn_l^T + nbc_a nbc_l^T + n_a n_l^T ;  t = c_true - R c_local
            T[0] = fma_(fa.n.x, l3.x, fma_(fa.nbc.x, l2.x, fa.bcn.x * l1.x));
            T[1] = fma_(fa.n.x, l3.y, fma_(fa.nbc.x, l2.y, fa.bcn.x * l1.y));
            T[2] = fma_(fa.n.x, l3.z, fma_(fa.nbc.x, l2.z, fa.bcn.x * l1.z));
            T[3] = fma_(fa.n.y, l3.x, fma_(fa.nbc.y, l2.x, fa.bcn.y * l1.x));
            T[4] = fma_(fa.n.y, l3.y, fma_(fa.nbc.y, l2.y, fa.bcn.y * l1.y));
            T[5] = fma_(fa.n.y, l3.z, fma_(fa.nbc.y, l2.z, fa.bcn.y * l1.z));
            T[6] = fma_(fa.n.z, l3.x, fma_(fa.nbc.z, l2.x, fa.bcn.z * l1.x));
            T[7] = fma_(fa.n.z, l3.y, fma_(fa.nbc.z, l2.y, fa.bcn.z * l1.y));
            T[8] = fma_(fa.n.z, l3.z, fma_(fa.nbc.z, l2.z, fa.bcn.z * l1.z));
            T[9] = c.x - fma_(T[2], lo.z, fma_(T[1], lo.y, T[0] * lo.x));
            T[10] = c.y - fma_(T[5], lo.z, fma_(T[4], lo.y, T[3] * lo.x));
            T[11] = c.z - fma_(T[8], lo.z, fma_(T[7], lo.y, T[6] * lo.x));
            t0 = xform(T, q0);
            t1 = xform(T, q1);
            t2 = xform(T, q2);
            for (int i = 0; i < 12; i++) SG(s, SEG_T + i) = T[i];
        }
        const float nf = (float)(3u * (a1 - a0 + 1u));  // atoms in the segment
        s0 = blend(t0, e0, 3.0f, nf - 3.0f, inv);
        s1 = blend(t1, e1, 2.0f, nf - 2.0f, inv);
        s2 = blend(t2, e2, 1.0f, nf - 1.0f, inv);
    }
    // blended tail of the last segment = final coordinates of the last residue (src/foldcomp.cpp:851-853);
    // parked in the S field of the closing anchor's slot, emitted by dec_blend
    ST3(n_seg, SEG_S, s0); ST3(n_seg, SEG_S + 3, s1); ST3(n_seg, SEG_S + 6, s2);
#undef SG
#undef LD3
#undef ST3
#undef SEG_S
#undef SEG_T
#undef SEG_TAIL
#undef SEG_F
#undef SEG_A
#undef SEG_I
#undef SEG_CS
}

template <class Ctx>
FCZ_HD void dec_stitch(Ctx& cx, const Tables* tb, const DecChain& ch) {
    (void)tb;
    if (cx.tid == 0) dec_stitch_core<SegFull>(ch.seg, 1, (int)ch.y.n_anchor - 1);
}

template <class Ctx>
FCZ_HD void dec_blend(Ctx& cx, const Tables* tb, const DecChain& ch) {
    const Layout& y = ch.y;
    const uint32_t L = y.L;
    const uint8_t* blob = ch.blob;
    const uint8_t* rec = blob + y.o_rec;
    const int n_seg = (int)y.n_anchor - 1;
    const uint32_t nT = 3u * L - 3u;
    (void)L; (void)rec; (void)n_seg; (void)nT; (void)blob; (void)tb;
    // ---- phase 4: blend (weightedAverage, src/atom_coordinate.cpp:145-163).
    //  (a) one lane per segment finishes the reverse pass: atoms 2,1,0 need the bond angles at the true
    //      atoms 3,2,1 (which involve the start atoms S) and emits the blended first residue;
    //  (b) every other backbone atom, one per thread: forward = local atom moved by T, reverse from phase 2.
    if (cx.tid == cx.nthr - 1) {  // last residue: the blended tail of the last segment (parked by the stitch)
        const float* sl = ch.seg + n_seg * FCZ_SEG_FLOATS + SEG_S;
        float* o = ch.out_xyz + 3u * ch.aoff[L - 1u];
        st3(o, ld3(sl)); st3(o + 3, ld3(sl + 3)); st3(o + 6, ld3(sl + 6));
    }
    for (int s = cx.tid; s < n_seg; s += cx.nthr) {
        const float* sg = ch.seg + s * FCZ_SEG_FLOATS;
        const float* T = sg + SEG_T;
        const uint32_t a0 = seg_a0(sg), a1 = seg_a1(sg);
        if (a1 <= a0) continue;  // empty segment: nothing to emit (its three atoms belong to the next one)
        const int n = (int)(3u * (a1 - a0 + 1u));
        const float inv = sg[SEG_I + 2];
        NerfFrame f;
        f.bcn = ld3(sg + SEG_RF); f.nbc = ld3(sg + SEG_RF + 3); f.n = ld3(sg + SEG_RF + 6);
        f3 rc = ld3(sg + SEG_RF + 9);
        // forward window: true atoms 3 and 4 (N', CA' of the first placed residue; the tail when n == 6)
        f3 f1 = xform(T, ld3(sg + SEG_HEAD)), f2 = xform(T, ld3(sg + SEG_HEAD + 3));
        float* slot = ch.out_xyz + 3u * ch.aoff[a0];
        for (int q = 2; q >= 0; q--) {
            const f3 f0 = ld3(sg + SEG_S + 3 * q);
            const cs ba = cossin_angle(f0, f1, f2);  // angle at true atom q+1
            const float bl = (q == 0) ? FCZ_N_TO_CA : (q == 1 ? FCZ_CA_TO_C : FCZ_C_TO_N);
            // torsion 3*a0+q: of the segment's first record (kept in the scratch by the forward lane)
            const cs tq = {sg[SEG_CS + 6 + 2 * q], sg[SEG_CS + 7 + 2 * q]};
            rc = nerf_step(f, rc, bl, ba, tq);
            st3(slot + 3 * q, blend(f0, rc, (float)(n - q), (float)q, inv));
            f2 = f1; f1 = f0;
        }
    }
    const float* locp = ch.loc ? ch.loc : nullptr;
#pragma unroll 3
    for (uint32_t g = 3u + cx.tid; g < 3u * L - 3u; g += cx.nthr) {
        const uint32_t r = g / 3u, k = g - 3u * r;
        // both inputs first (global memory on the three-kernel path): their latency overlaps across the unrolled trips
        const f3 rv = ld3(ch.rev + 3u * g);
        float* slot = ch.out_xyz + 3u * (ch.aoff[r] + k);
        const f3 local = locp ? ld3(locp + 3u * g) : ld3(slot);
        const int s = ch.segid[r - 1u];  // residue r was placed while consuming record r-1
        const float* sg = ch.seg + s * FCZ_SEG_FLOATS;
        const uint32_t a0 = seg_a0(sg), a1 = seg_a1(sg);
        if (r == a0 || r >= a1) {
            // first residue of the NEXT segment (emitted by (a)) -- or the chain's last residue (phase 3)
            continue;
        }
        const int q = (int)(g - 3u * a0), n = (int)(3u * (a1 - a0 + 1u));
        st3(slot, blend(xform(sg + SEG_T, local), rv, (float)(n - q), (float)q, sg[SEG_I + 2]));
    }
}

template <class Ctx>
FCZ_HD void dec_side(Ctx& cx, const Tables* tb, const DecChain& ch) {
    const Layout& y = ch.y;
    const uint32_t L = y.L;
    const uint8_t* blob = ch.blob;
    const uint8_t* rec = blob + y.o_rec;
    const int n_seg = (int)y.n_anchor - 1;
    const uint32_t nT = 3u * L - 3u;
    (void)L; (void)rec; (void)n_seg; (void)nT; (void)blob; (void)tb;
    // ---- phase 5: side chains, one thread per PAIR of neighbouring residues (two independent dependency
    // chains per thread hide the latency of a placement; no barrier is needed because an atom's predecessors
    // are lower slots of the same residue).  Nerf::reconstructAminoAcid src/nerf.cpp:106-155; torsion =
    // FixedAngleDiscretizer(255).continuize(byte), src/foldcomp.cpp:338-369, read from a 256-entry table.
    {
        const uint8_t* sc = ch.sc ? ch.sc : blob + y.o_sc;
        // Residues sorted by atom count (counting sort, longest first) so that the lanes of a warp -- and the two
        // residues of a pair -- run the same number of placements: a warp costs its longest lane.
        const uint16_t* ord = nullptr;
        if (ch.order) {
            uint32_t* bins = ch.bins;  // [0,16) counts by atom count, [16,32) write cursors
            for (uint32_t i = cx.tid; i < 32u; i += cx.nthr) bins[i] = 0u;
            cx.sync();
            for (uint32_t r = cx.tid; r < L; r += cx.nthr) cx.atomic_add(&bins[(ch.aoff[r + 1u] - ch.aoff[r]) & 15u], 1u);
            cx.sync();
            if (cx.tid == 0) {
                uint32_t acc = 0;
                for (int k = 15; k >= 0; k--) { bins[16 + k] = acc; acc += bins[k]; }
            }
            cx.sync();
            for (uint32_t r = cx.tid; r < L; r += cx.nthr) ch.order[cx.atomic_add(&bins[16u + ((ch.aoff[r + 1u] - ch.aoff[r]) & 15u)], 1u)] = (uint16_t)r;
            cx.sync();
            ord = ch.order;
        }
        // one lane per residue; warp-sized runs of the sorted order are dealt to the warps boustrophedon
        // (0..n-1, n-1..0, ...) so that every warp gets long and short runs
        const uint32_t runs = (L + cx.wsize - 1u) / cx.wsize;
        for (uint32_t cyc = 0; cyc * cx.nwarps < runs; cyc++) {
            const uint32_t j = cyc * cx.nwarps + ((cyc & 1u) ? (uint32_t)(cx.nwarps - 1 - cx.warp) : (uint32_t)cx.warp);
            const uint32_t i = j * cx.wsize + cx.lane;
            if (j >= runs || i >= L) continue;
            const uint32_t r = ord ? ord[i] : i;
            const unsigned code = ch.codes ? ch.codes[r] : (unsigned)(rec[8u * r] >> 3);
            const uint32_t o = ch.aoff[r], n = ch.aoff[r + 1u] - o;
            float* R = ch.out_xyz + 3u * o;
            const uint8_t* sb = sc + (o - 3u * r);
            // the table entries of placement k+1 are fetched while placement k is computed (they depend on the residue code
            // and the stored byte only): the dependent look-ups were the kernel's hottest line
            const Tables::ScEntry* ent = tb->sc_ent[code];
            Tables::ScEntry e_n = ent[3];
            cs tor_n = {0.f, 0.f};
            if (n > 3u) tor_n = tb->sc_tor[sb[0]];
            for (uint32_t k = 3u; k < n; k++) {
                const Tables::ScEntry e = e_n;
                const cs tor = tor_n;
                if (k + 1u < n) { e_n = ent[k + 1u]; tor_n = tb->sc_tor[sb[k - 2u]]; }
                const unsigned pp = e.pred;
                st3(R + 3u * k, place_from(ld3(R + 3u * (pp & 15u)), ld3(R + 3u * ((pp >> 4) & 15u)), ld3(R + 3u * ((pp >> 8) & 15u)), e.blen, e.bang, tor));
            }
        }
        cx.sync();
        if (ch.use_alt) {  // _reorderAtoms, src/foldcomp.cpp:1563-1577
            for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
                const unsigned code = rec[8u * r] >> 3;
                const uint32_t na = tb->natoms[code];
                float* R = ch.out_xyz + 3u * ch.aoff[r];
                f3 tmp[FCZ_MAX_ATOMS];
                for (uint32_t k = 0; k < na; k++) tmp[k] = ld3(R + 3u * k);
                for (uint32_t k = 0; k < na; k++) st3(R + 3u * k, tmp[tb->alt[code][k]]);
            }
        }
    }
}

template <class Ctx>
FCZ_HD void decode_chain(Ctx& cx, const Tables* tb, const DecChain& ch) {
    dec_unpack(cx, tb, ch);
    cx.sync();
    cx.mark(8);  // D_UNPACK
    dec_passes(cx, tb, ch);
    cx.sync();
    cx.mark(9);  // D_PASSES
    dec_stitch(cx, tb, ch);
    cx.sync();
    cx.mark(10);  // D_STITCH
    dec_blend(cx, tb, ch);
    cx.sync();
    cx.mark(11);  // D_BLEND
    dec_side(cx, tb, ch);
    cx.sync();
    cx.mark(12);  // D_SIDE
}

}  // namespace fcz
#endif  // FCZ_CODEC_H
