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

#include "../../include/fcz_engine.h"
#include "fcz_format.h"

namespace fcz {

// Device-friendly copy of the residue tables (built once from fcz_tables.h).
struct Tables {
    uint8_t natoms[FCZ_NUM_CODES];
    uint8_t name1[FCZ_NUM_CODES];  // one-letter codes (header firstResidue / lastResidue)
    uint8_t alt[FCZ_NUM_CODES][FCZ_MAX_ATOMS];
    uint16_t pred[FCZ_NUM_CODES][FCZ_MAX_ATOMS];
    float blen[FCZ_NUM_CODES][FCZ_MAX_ATOMS];
    cs bang[FCZ_NUM_CODES][FCZ_MAX_ATOMS];  // (cos, sin) of the table bond angle
};

// Filled on the host with host libm: the table angles are compile-time constants of the format,
// so their (cos,sin) are computed once, exactly as cossin_deg would.
inline void build_tables(Tables* t) {
    for (int c = 0; c < FCZ_NUM_CODES; c++) {
        t->natoms[c] = FCZ_NATOMS[c];
        t->name1[c] = (uint8_t)FCZ_NAME1[c];
        for (int k = 0; k < FCZ_MAX_ATOMS; k++) {
            t->alt[c][k] = FCZ_ALT[c][k];
            t->pred[c][k] = FCZ_PRED[c][k];
            t->blen[c][k] = FCZ_BLEN[c][k];
            float r = (float)((double)FCZ_BANG[c][k] * M_PI / 180.0);
            t->bang[c][k].c = cosf(r);
            t->bang[c][k].s = sinf(r);
        }
    }
}

// ------------------------------------------------------------------------------------------ encode

// scratch floats needed by the min/max reduction: per-warp partials + results + discretiser params
#define FCZ_RED_FLOATS(nwarps) ((nwarps) * 14 + 14 + 21)

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
    uint16_t* ares;          // [A]   residue of each atom
    float* ang;              // [6*L] the six backbone arrays, header order, stride L
    float* red;              // [FCZ_RED_FLOATS(nwarps)]
};

FCZ_HD f3 bb_atom(const EncChain& ch, uint32_t j) {  // j-th backbone atom (N,CA,C = slots 0..2)
    uint32_t r = j / 3u, k = j - 3u * r;
    return ld3(ch.X + 3u * (ch.aoff[r] + k));
}

template <class Ctx>
FCZ_HD void encode_chain(Ctx& cx, const Tables* tb, const EncChain& ch) {
    const uint32_t L = ch.L, A = ch.A;
    const int n_anchor = anchor_count(L, ch.b);
    const Layout y = make_layout(L, A - 3u * L, ch.title_len, (uint32_t)n_anchor);
    uint8_t* B = ch.B;

    // ---- phase 1: residue -> first atom (exclusive scan of table atom counts), atom -> residue
    {
        const uint32_t chunk = (L + cx.nthr - 1) / cx.nthr;
        uint32_t r0 = cx.tid * chunk; if (r0 > L) r0 = L;
        uint32_t r1 = r0 + chunk; if (r1 > L) r1 = L;
        uint32_t sum = 0;
        for (uint32_t r = r0; r < r1; r++) sum += tb->natoms[ch.type[r]];
        uint32_t base = cx.excl_scan(sum);
        for (uint32_t r = r0; r < r1; r++) {
            ch.aoff[r] = base;
            uint32_t n = tb->natoms[ch.type[r]];
            for (uint32_t k = 0; k < n; k++) ch.ares[base + k] = (uint16_t)r;
            base += n;
        }
        if (r1 == L) ch.aoff[L] = base;
    }
    cx.stage_wait();  // coordinates staged by the caller are now visible
    cx.sync();

    // ---- phase 2: one dihedral per atom.  Side-chain atoms (slot >= 3) give the side-chain byte
    // (src/sidechain.cpp:149-168 + FixedAngleDiscretizer, src/foldcomp.cpp:532-538); backbone
    // atom j = 3r+k gives backbone torsion j (src/torsion_angle.cpp:49-94): k=0 psi, 1 omega, 2 phi
    // (src/foldcomp.cpp:488-492).
    {
        const float mn = sc_min(), df = sc_disc_f();
        const uint32_t nT = 3u * L - 3u;
        for (uint32_t a = cx.tid; a < A; a += cx.nthr) {
            uint32_t r = ch.ares[a];
            uint32_t a0 = ch.aoff[r];
            uint32_t k = a - a0;
            if (k >= 3u) {
                unsigned pr = tb->pred[ch.type[r]][k];
                const float* R = ch.X + 3u * a0;
                float t = dihedral_deg(ld3(R + 3u * (pr & 15u)), ld3(R + 3u * ((pr >> 4) & 15u)),
                                       ld3(R + 3u * ((pr >> 8) & 15u)), ld3(R + 3u * k));
                B[y.o_sc + (a - 3u * (r + 1u))] = (uint8_t)disc_trunc(t, mn, df);
            } else {
                uint32_t j = 3u * r + k;
                if (j < nT) {
                    float t = dihedral_deg(bb_atom(ch, j), bb_atom(ch, j + 1), bb_atom(ch, j + 2), bb_atom(ch, j + 3));
                    int arr = (k == 0) ? A_PSI : (k == 1 ? A_OMEGA : A_PHI);
                    ch.ang[arr * L + r] = t;
                }
            }
        }
    }
    // ---- phase 3: backbone bond angles at atoms m = 2 .. 3L-2 (src/nerf.cpp:495-508; split by
    // index%3 at src/foldcomp.cpp:496-505: m%3==2 CA-C-N[m/3], 0 C-N-CA[m/3-1], 1 N-CA-C[m/3-1])
    for (uint32_t m = 2u + cx.tid; m + 1u < 3u * L; m += cx.nthr) {
        float v = bond_angle_deg(bb_atom(ch, m - 1), bb_atom(ch, m), bb_atom(ch, m + 1));
        uint32_t q = m / 3u, k = m - 3u * q;
        if (k == 2u) ch.ang[A_CACN * L + q] = v;
        else if (k == 0u) ch.ang[A_CNCA * L + q - 1u] = v;
        else ch.ang[A_NCAC * L + q - 1u] = v;
    }
    cx.sync();

    // ---- phase 4: min / max of the six arrays (L-1 values) and of the B-factors (L values)
    // (Discretizer::Discretizer, src/discretizer.cpp:22-33)
    {
        float mn[7], mx[7];
        for (int k = 0; k < 7; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
        for (uint32_t i = cx.tid; i < L; i += cx.nthr) {
            if (i + 1u < L) {
                for (int k = 0; k < 6; k++) {
                    float v = ch.ang[k * L + i];
                    mn[k] = min_ignore_nan(mn[k], v);
                    mx[k] = max_ignore_nan(mx[k], v);
                }
            }
            float v = ch.bfac[i];
            mn[6] = min_ignore_nan(mn[6], v);
            mx[6] = max_ignore_nan(mx[6], v);
        }
        for (int k = 0; k < 7; k++) {
            mn[k] = cx.wmin(mn[k]);
            mx[k] = cx.wmax(mx[k]);
        }
        if (cx.lane == 0) {
            for (int k = 0; k < 7; k++) {
                ch.red[cx.warp * 14 + k] = mn[k];
                ch.red[cx.warp * 14 + 7 + k] = mx[k];
            }
        }
        cx.sync();
        float* res = ch.red + cx.nwarps * 14;  // [14] min/max, then [21] (min, disc_f, cont_f) x 7
        for (int k = cx.tid; k < 7; k += cx.nthr) {
            float lo = INFINITY, hi = -INFINITY;
            for (int w = 0; w < cx.nwarps; w++) {
                lo = min_ignore_nan(lo, ch.red[w * 14 + k]);
                hi = max_ignore_nan(hi, ch.red[w * 14 + 7 + k]);
            }
            float first = (k < 6) ? ch.ang[k * L] : ch.bfac[0];
            if (first != first) { lo = first; hi = first; }  // min_element/max_element keep a NaN first element
            unsigned nb = (k < 6) ? n_bins(k) : 255u;
            float* prm = res + 14 + 3 * k;
            prm[0] = lo;
            prm[1] = disc_factor(lo, hi, nb);
            prm[2] = cont_factor(lo, hi, nb);
        }
        cx.sync();
    }
    const float* prm = ch.red + cx.nwarps * 14 + 14;

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
    for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
        unsigned q[6] = {0, 0, 0, 0, 0, 0};
        if (r + 1u < L) {
            for (int k = 0; k < 6; k++) q[k] = disc_round(ch.ang[k * L + r], prm[3 * k], prm[3 * k + 1]);
        }
        pack_record(B + y.o_rec + 8u * r, ch.type[r], q[A_PHI], q[A_PSI], q[A_OMEGA], q[A_NCAC], q[A_CACN], q[A_CNCA]);
        B[y.o_temp + 8u + r] = (uint8_t)disc_round(ch.bfac[r], prm[18], prm[19]);
    }
    cx.sync();
}

// ------------------------------------------------------------------------------------------ decode

// floats of per-segment scratch: S[9] actual start atoms, T[12] rigid transform, tail[9] local tail,
// F[12] local frame (3 axes + origin)
#define FCZ_SEG_FLOATS 42
enum { SEG_S = 0, SEG_T = 9, SEG_TAIL = 21, SEG_F = 30 };

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
    float* seg;           // [(n_anchor-1) * FCZ_SEG_FLOATS]
};

// orthonormal frame of a triangle (p0,p1,p2): e1 along p0->p1, e3 normal, e2 = e3 x e1
struct Frame {
    f3 e1, e2, e3;
};
FCZ_HD f3 scale3(f3 v, float s) { return mk3(v.x * s, v.y * s, v.z * s); }
FCZ_HD float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
FCZ_HD Frame make_frame(f3 p0, f3 p1, f3 p2) {
    Frame f;
    f3 u = sub3(p1, p0);
    f.e1 = scale3(u, 1.0f / norm3(u));
    f3 n = cross3(f.e1, sub3(p2, p1));
    f.e3 = scale3(n, 1.0f / norm3(n));
    f.e2 = cross3(f.e3, f.e1);
    return f;
}
// x -> R (x - o_loc) + o_act with R = F_act F_loc^T, evaluated as sum_j e_act_j * (e_loc_j . (x - o_loc))
FCZ_HD f3 xform(const float* T, f3 x) {
    // T: rows of R (9), then t (3):  y = R x + t
    return mk3(((T[0] * x.x + T[1] * x.y) + T[2] * x.z) + T[9], ((T[3] * x.x + T[4] * x.y) + T[5] * x.z) + T[10],
               ((T[6] * x.x + T[7] * x.y) + T[8] * x.z) + T[11]);
}

FCZ_HD float n_ca_len(unsigned code) { return code == FCZ_CODE_PRO ? FCZ_PRO_N_TO_CA : FCZ_N_TO_CA; }

template <class Ctx>
FCZ_HD void decode_chain(Ctx& cx, const Tables* tb, const DecChain& ch) {
    const Layout& y = ch.y;
    const uint32_t L = y.L;
    const uint8_t* blob = ch.blob;
    const uint8_t* rec = blob + y.o_rec;
    const int n_seg = (int)y.n_anchor - 1;
    const uint32_t nT = 3u * L - 3u;

    // ---- phase 1: records -> residue codes, atom offsets (exclusive scan), (cos,sin) of the
    // continuised angles (convertBytesToBackboneChain src/foldcomp.cpp:60-77, decompressBackboneChain
    // 122-153, _continuize 155-158; the deg->rad and sincos of Nerf::place_atom src/nerf.cpp:63-70
    // are hoisted here so the recurrences below carry no transcendental)
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
        for (uint32_t r = r0; r < r1; r++) sum += tb->natoms[rec[8u * r] >> 3];
        uint32_t base = cx.excl_scan(sum);
        for (uint32_t r = r0; r < r1; r++) {
            ch.aoff[r] = base;
            base += tb->natoms[rec[8u * r] >> 3];
        }
        if (r1 == L) ch.aoff[L] = base;
        const float tmin = get_f32(blob + y.o_temp), tcf = get_f32(blob + y.o_temp + 4);
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
            Record q = unpack_record(rec + 8u * r);
            ch.out_type[r] = (uint8_t)q.res;
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
    cx.sync();

    // ---- phase 2: forward NeRF pass of every anchor segment, one lane per segment, in the
    // segment's LOCAL frame: it starts from the STORED anchor instead of the blended tail of the
    // previous segment (which is not known yet).  From its 4th placed atom on a forward pass is a
    // rigid body hanging off its first placed residue (N',CA',C'), so the true pass is this one
    // moved by a rigid transform that phase 3 determines.  (reconstructBackboneAtoms,
    // src/foldcomp.cpp:167-246; Pro N-CA length taken from the record being consumed, 204-212.)
    for (int s = cx.tid; s < n_seg; s += cx.nthr) {
        float* sg = ch.seg + s * FCZ_SEG_FLOATS;
        const uint32_t a0 = get_u32(blob + y.o_aidx + 4u * s), a1 = get_u32(blob + y.o_aidx + 4u * (s + 1));
        const uint8_t* anc = blob + y.o_anchor + 36u * s;
        f3 p0 = mk3(get_f32(anc), get_f32(anc + 4), get_f32(anc + 8));
        f3 p1 = mk3(get_f32(anc + 12), get_f32(anc + 16), get_f32(anc + 20));
        f3 p2 = mk3(get_f32(anc + 24), get_f32(anc + 28), get_f32(anc + 32));
        for (uint32_t r = a0; r < a1; r++) {
            const uint32_t t = 3u * r;
            f3 n = place_atom(p0, p1, p2, FCZ_C_TO_N, ch.ang[t], ch.tor[t]);
            f3 ca = place_atom(p1, p2, n, n_ca_len(rec[8u * r] >> 3), ch.ang[t + 1u], ch.tor[t + 1u]);
            f3 c = place_atom(p2, n, ca, FCZ_CA_TO_C, ch.ang[t + 2u], ch.tor[t + 2u]);
            float* o = ch.out_xyz + 3u * ch.aoff[r + 1u];
            st3(o, n); st3(o + 3, ca); st3(o + 6, c);
            if (r == a0) {  // local frame of the first placed residue
                Frame f = make_frame(n, ca, c);
                st3(sg + SEG_F, f.e1); st3(sg + SEG_F + 3, f.e2); st3(sg + SEG_F + 6, f.e3); st3(sg + SEG_F + 9, n);
            }
            p0 = n; p1 = ca; p2 = c;
        }
        st3(sg + SEG_TAIL, p0); st3(sg + SEG_TAIL + 3, p1); st3(sg + SEG_TAIL + 6, p2);
    }
    cx.sync();

    // ---- phase 3: stitch.  Serial over segments (the only cross-segment dependency of the
    // reference, src/foldcomp.cpp:855-857: the blended tail of segment s seeds segment s+1).  Per
    // segment: place N',CA',C' from the true start atoms, derive the rigid transform local->true,
    // move the local tail, blend it with the stored anchor (weightedAverage,
    // src/atom_coordinate.cpp:145-163, last three atoms only).
    if (cx.tid == 0) {
        f3 s0, s1, s2;
        {
            const uint8_t* anc = blob + y.o_anchor;
            s0 = mk3(get_f32(anc), get_f32(anc + 4), get_f32(anc + 8));
            s1 = mk3(get_f32(anc + 12), get_f32(anc + 16), get_f32(anc + 20));
            s2 = mk3(get_f32(anc + 24), get_f32(anc + 28), get_f32(anc + 32));
        }
        for (int s = 0; s < n_seg; s++) {
            float* sg = ch.seg + s * FCZ_SEG_FLOATS;
            st3(sg + SEG_S, s0); st3(sg + SEG_S + 3, s1); st3(sg + SEG_S + 6, s2);
            const uint32_t a0 = get_u32(blob + y.o_aidx + 4u * s), a1 = get_u32(blob + y.o_aidx + 4u * (s + 1));
            const uint8_t* anc = blob + y.o_anchor + 36u * (s + 1);
            f3 e0 = mk3(get_f32(anc), get_f32(anc + 4), get_f32(anc + 8));
            f3 e1 = mk3(get_f32(anc + 12), get_f32(anc + 16), get_f32(anc + 20));
            f3 e2 = mk3(get_f32(anc + 24), get_f32(anc + 28), get_f32(anc + 32));
            f3 t0 = s0, t1 = s1, t2 = s2;  // forward tail in true coordinates
            if (a1 > a0) {
                const uint32_t t = 3u * a0;
                f3 n = place_atom(s0, s1, s2, FCZ_C_TO_N, ch.ang[t], ch.tor[t]);
                f3 ca = place_atom(s1, s2, n, n_ca_len(rec[8u * a0] >> 3), ch.ang[t + 1u], ch.tor[t + 1u]);
                f3 c = place_atom(s2, n, ca, FCZ_CA_TO_C, ch.ang[t + 2u], ch.tor[t + 2u]);
                Frame fa = make_frame(n, ca, c);
                f3 l1 = ld3(sg + SEG_F), l2 = ld3(sg + SEG_F + 3), l3 = ld3(sg + SEG_F + 6), lo = ld3(sg + SEG_F + 9);
                float* T = sg + SEG_T;
                // R = e_act1 l1^T + e_act2 l2^T + e_act3 l3^T
                T[0] = (fa.e1.x * l1.x + fa.e2.x * l2.x) + fa.e3.x * l3.x;
                T[1] = (fa.e1.x * l1.y + fa.e2.x * l2.y) + fa.e3.x * l3.y;
                T[2] = (fa.e1.x * l1.z + fa.e2.x * l2.z) + fa.e3.x * l3.z;
                T[3] = (fa.e1.y * l1.x + fa.e2.y * l2.x) + fa.e3.y * l3.x;
                T[4] = (fa.e1.y * l1.y + fa.e2.y * l2.y) + fa.e3.y * l3.y;
                T[5] = (fa.e1.y * l1.z + fa.e2.y * l2.z) + fa.e3.y * l3.z;
                T[6] = (fa.e1.z * l1.x + fa.e2.z * l2.x) + fa.e3.z * l3.x;
                T[7] = (fa.e1.z * l1.y + fa.e2.z * l2.y) + fa.e3.z * l3.y;
                T[8] = (fa.e1.z * l1.z + fa.e2.z * l2.z) + fa.e3.z * l3.z;
                // t = o_act - R o_loc
                T[9] = n.x - ((T[0] * lo.x + T[1] * lo.y) + T[2] * lo.z);
                T[10] = n.y - ((T[3] * lo.x + T[4] * lo.y) + T[5] * lo.z);
                T[11] = n.z - ((T[6] * lo.x + T[7] * lo.y) + T[8] * lo.z);
                t0 = xform(T, ld3(sg + SEG_TAIL));
                t1 = xform(T, ld3(sg + SEG_TAIL + 3));
                t2 = xform(T, ld3(sg + SEG_TAIL + 6));
            }
            const float nf = (float)(3u * (a1 - a0 + 1u));  // atoms in the segment
            const float w0 = nf - 3.0f, w1 = nf - 2.0f, w2 = nf - 1.0f;  // index i of the tail atoms
            s0 = mk3((t0.x * 3.0f + e0.x * w0) / nf, (t0.y * 3.0f + e0.y * w0) / nf, (t0.z * 3.0f + e0.z * w0) / nf);
            s1 = mk3((t1.x * 2.0f + e1.x * w1) / nf, (t1.y * 2.0f + e1.y * w1) / nf, (t1.z * 2.0f + e1.z * w1) / nf);
            s2 = mk3((t2.x * 1.0f + e2.x * w2) / nf, (t2.y * 1.0f + e2.y * w2) / nf, (t2.z * 1.0f + e2.z * w2) / nf);
        }
        // blended tail of the last segment = final coordinates of the last residue (src/foldcomp.cpp:851-853)
        float* o = ch.out_xyz + 3u * ch.aoff[L - 1u];
        st3(o, s0); st3(o + 3, s1); st3(o + 6, s2);
    }
    cx.sync();

    // ---- phase 4: reverse pass + blend, one lane per segment (reconstructBackboneReverse,
    // src/foldcomp.cpp:248-273; Nerf::reconstructWithReversed src/nerf.cpp:342-379 with forward
    // indices; bond angles recomputed from the FORWARD atoms as getBondAngles src/nerf.cpp:495-508
    // does; bond lengths by atom kind, never the Pro length, src/nerf.h:37-43).
    for (int s = cx.tid; s < n_seg; s += cx.nthr) {
        const float* sg = ch.seg + s * FCZ_SEG_FLOATS;
        const float* T = sg + SEG_T;
        const uint32_t a0 = get_u32(blob + y.o_aidx + 4u * s), a1 = get_u32(blob + y.o_aidx + 4u * (s + 1));
        if (a1 <= a0) continue;  // empty segment: nothing to emit (its three atoms belong to the next one)
        const int n = (int)(3u * (a1 - a0 + 1u));
        const float nf = (float)n;
        const uint8_t* anc = blob + y.o_anchor + 36u * (s + 1);
        // reversed chain window (atoms q+3, q+2, q+1) starts as the stored anchor C, CA, N
        f3 r1 = mk3(get_f32(anc), get_f32(anc + 4), get_f32(anc + 8));
        f3 r2 = mk3(get_f32(anc + 12), get_f32(anc + 16), get_f32(anc + 20));
        f3 r3 = mk3(get_f32(anc + 24), get_f32(anc + 28), get_f32(anc + 32));
        // forward window (atoms q+1, q+2) starts as the moved local tail
        f3 f1 = xform(T, ld3(sg + SEG_TAIL)), f2 = xform(T, ld3(sg + SEG_TAIL + 3));
        for (int q = n - 4; q >= 0; q--) {
            const uint32_t r = a0 + (uint32_t)q / 3u, k = (uint32_t)q % 3u;
            float* slot = ch.out_xyz + 3u * (ch.aoff[r] + k);
            f3 f0 = (q < 3) ? ld3(sg + SEG_S + 3 * q) : xform(T, ld3(slot));
            cs ba = cossin_deg(bond_angle_deg(f0, f1, f2));  // angle at forward atom q+1
            const float bl = (k == 0u) ? FCZ_N_TO_CA : (k == 1u ? FCZ_CA_TO_C : FCZ_C_TO_N);
            uint32_t ti = 3u * a0 + (uint32_t)q;
            if (ti >= nT) ti = nT - 1u;
            f3 nw = place_atom(r3, r2, r1, bl, ba, ch.tor[ti]);
            const float wf = (float)(n - q), wr = (float)q;
            st3(slot, mk3((f0.x * wf + nw.x * wr) / nf, (f0.y * wf + nw.y * wr) / nf, (f0.z * wf + nw.z * wr) / nf));
            r3 = r2; r2 = r1; r1 = nw;
            f2 = f1; f1 = f0;
        }
    }
    cx.sync();

    // ---- phase 5: side chains, one lane per residue (Nerf::reconstructAminoAcid
    // src/nerf.cpp:106-155; torsion = FixedAngleDiscretizer(255).continuize(byte),
    // src/foldcomp.cpp:338-369).  Atoms are built in place in the output area: predecessors of an
    // atom always have lower slots in the same residue.
    {
        const float mn = sc_min(), cf = sc_cont_f();
        const uint8_t* sc = blob + y.o_sc;
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) {
            const unsigned code = rec[8u * r] >> 3;
            const uint32_t a0 = ch.aoff[r];
            const uint32_t na = tb->natoms[code];
            float* R = ch.out_xyz + 3u * a0;
            const uint8_t* tb_sc = sc + (a0 - 3u * r);
            for (uint32_t k = 3u; k < na; k++) {
                unsigned pr = tb->pred[code][k];
                cs to = cossin_deg(continuize(tb_sc[k - 3u], mn, cf));
                f3 v = place_atom(ld3(R + 3u * (pr & 15u)), ld3(R + 3u * ((pr >> 4) & 15u)),
                                  ld3(R + 3u * ((pr >> 8) & 15u)), tb->blen[code][k], tb->bang[code][k], to);
                st3(R + 3u * k, v);
            }
            if (ch.use_alt) {  // _reorderAtoms, src/foldcomp.cpp:1563-1577
                f3 tmp[FCZ_MAX_ATOMS];
                for (uint32_t k = 0; k < na; k++) tmp[k] = ld3(R + 3u * k);
                for (uint32_t k = 0; k < na; k++) st3(R + 3u * k, tmp[tb->alt[code][k]]);
            }
        }
    }
    cx.sync();
}

}  // namespace fcz
#endif  // FCZ_CODEC_H
