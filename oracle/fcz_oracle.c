/* oracle/fcz_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See fcz_oracle.h.
 *
 * Scalar CPU restatement of the reference's FCZ codec.  Every function cites the reference
 * file:line (relative to /root/reference/) whose arithmetic and operation ORDER it follows; the
 * float/double rounding contract is the one in SURVEY.md Appendix B.  Build with the pinned flags
 * of oracle/Makefile (-O3 -ffp-contract=off, x86-64 baseline, glibc libm) so that no FMA is formed.
 *
 * Data layout differences from the reference are deliberate (integer slot tables instead of
 * string-keyed maps, flat arrays instead of vector<AtomCoordinate>); the arithmetic is not.
 */
#include "fcz_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../foldcomp_b200/csrc/fcz_tables.h"

typedef struct { float x, y, z; } f3;

/* ---------------------------------------------------------------- vector math (src/float3d.h) */

/* src/float3d.h:19-25 crossProduct: float mul/sub, no FMA */
static f3 cross(f3 a, f3 b) {
    f3 r;
    r.x = a.y * b.z - b.y * a.z;
    r.y = a.z * b.x - b.z * a.x;
    r.z = a.x * b.y - b.x * a.y;
    return r;
}

static f3 sub(f3 a, f3 b) { f3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }

/* src/float3d.h:33-35 norm: pow(float,2) promotes to double, sum and sqrt in double, result float */
static float norm3(f3 v) {
    double s = (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z;
    return (float)sqrt(s);
}

/* src/float3d.h:36-43 getCosineTheta: float dot and squared sizes, float product of the sizes,
 * then DOUBLE sqrt and DOUBLE divide (unqualified sqrt on a float resolves to ::sqrt(double)),
 * rounded to float on assignment. */
static float cos_theta(f3 v1, f3 v2) {
    float inner = (v1.x * v2.x) + (v1.y * v2.y) + (v1.z * v2.z);
    float s1 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    float s2 = v2.x * v2.x + v2.y * v2.y + v2.z * v2.z;
    return (float)((double)inner / sqrt((double)(s1 * s2)));
}

/* src/float3d.h:55-65 angle: bond angle at atm2 in degrees; double acos, *180.0/M_PI in double */
static float angle3(f3 a1, f3 a2, f3 a3) {
    f3 d1 = sub(a1, a2), d2 = sub(a3, a2);
    float c = cos_theta(d1, d2);
    return (float)(acos((double)c) * 180.0 / M_PI);
}

/* src/torsion_angle.cpp:49-94 getTorsionFromXYZ, one dihedral */
static float dihedral(f3 a1, f3 a2, f3 a3, f3 a4) {
    f3 d1 = sub(a2, a1), d2 = sub(a3, a2), d3 = sub(a4, a3);
    f3 u1 = cross(d1, d2), u2 = cross(d2, d3);
    float c = cos_theta(u1, u2);
    float t;
    double ac = acos((double)c);
    if (isnan(ac)) {
        t = (c < 0) ? 180.0f : 0.0f; /* torsion_angle.cpp:74-79 */
    } else {
        t = (float)(ac * 180.0 / M_PI);
    }
    f3 pb = cross(u2, d2); /* torsion_angle.cpp:87-92 sign */
    if ((u1.x * pb.x) + (u1.y * pb.y) + (u1.z * pb.z) < 0) t = -1 * t;
    return t;
}

/* src/nerf.cpp:39-104 Nerf::place_atom */
static f3 place_atom(f3 a, f3 b, f3 c, float bond_length, float bond_angle, float torsion_angle) {
    f3 ab = sub(b, a), bc = sub(c, b);
    float bc_norm = norm3(bc);
    f3 bcn = {bc.x / bc_norm, bc.y / bc_norm, bc.z / bc_norm};
    bond_angle = (float)((double)bond_angle * M_PI / 180.0);       /* nerf.cpp:63 */
    torsion_angle = (float)((double)torsion_angle * M_PI / 180.0); /* nerf.cpp:64 */
    f3 cur;
    cur.x = (-1 * bond_length) * cosf(bond_angle);
    cur.y = (bond_length * cosf(torsion_angle)) * sinf(bond_angle);
    cur.z = (bond_length * sinf(torsion_angle)) * sinf(bond_angle);
    f3 n = cross(ab, bcn);
    float n_norm = norm3(n);
    n.x = n.x / n_norm; n.y = n.y / n_norm; n.z = n.z / n_norm;
    f3 nbc = cross(n, bcn);
    f3 d = {0.0f, 0.0f, 0.0f}; /* nerf.cpp:86-98: m = [bcn nbc n], accumulated term by term */
    d.x += bcn.x * cur.x; d.x += nbc.x * cur.y; d.x += n.x * cur.z;
    d.y += bcn.y * cur.x; d.y += nbc.y * cur.y; d.y += n.y * cur.z;
    d.z += bcn.z * cur.x; d.z += nbc.z * cur.y; d.z += n.z * cur.z;
    d.x += c.x; d.y += c.y; d.z += c.z;
    return d;
}

/* -------------------------------------------------------------- discretiser (src/discretizer.cpp) */

typedef struct { float min, max, disc_f, cont_f; } disc_t;

/* src/discretizer.cpp:22-33: std::min_element / max_element semantics (first element kept when
 * comparisons with NaN are false), float factors with n_bin converted to float. */
static disc_t disc_fit(const float* v, uint32_t n, unsigned nb) {
    disc_t d;
    float mn = v[0], mx = v[0];
    for (uint32_t i = 1; i < n; i++) {
        if (v[i] < mn) mn = v[i];
        if (mx < v[i]) mx = v[i];
    }
    d.min = mn; d.max = mx;
    d.disc_f = (float)nb / (mx - mn);
    d.cont_f = (mx - mn) / (float)nb;
    return d;
}

/* x86-64 double->unsigned as gcc emits it: cvttsd2si to int64, keep the low 32 bits; NaN and
 * out-of-range give 0x8000000000000000 -> 0 (src/discretizer.cpp:49 relies on this for constant
 * arrays where disc_f = inf). */
static unsigned d2u(double x) {
    if (!(x > -9.2e18 && x < 9.2e18)) return 0u;
    return (unsigned)(int64_t)x;
}
/* src/discretizer.cpp:49: (unsigned)((x - min) * disc_f + 0.5), the +0.5 in double */
static unsigned disc_round(const disc_t* d, float x) { return d2u((double)((x - d->min) * d->disc_f) + 0.5); }
/* src/discretizer.cpp:55-57: truncating scalar variant used for side chains (foldcomp.cpp:532-538) */
static unsigned disc_trunc(float min, float disc_f, float x) { return d2u((double)((x - min) * disc_f)); }
/* src/discretizer.cpp:64,71 / foldcomp.cpp:155-158: q*cont_f + min, two float roundings */
static float cont(unsigned q, float min, float cont_f) { return ((float)q * cont_f) + min; }

/* ------------------------------------------------------------------------------- format helpers */

#define FCZ_HDR 76 /* magic 4 + CompressedFileHeader 72 (src/foldcomp.h:118-136) */

static void put16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }
static void put32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static void putf(uint8_t* p, float v) { memcpy(p, &v, 4); }
static uint16_t get16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t get32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static float getf(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

/* src/foldcomp.cpp:739-761 _getAnchorNum/_setAnchor */
static int anchors(uint32_t L, int32_t b, int32_t* idx /* may be NULL */) {
    int n_inner = (int)L / b;
    int n_all = n_inner + 2;
    int interval = (int)L / (n_all - 1);
    if (idx) {
        for (int i = 0; i < n_all - 1; i++) idx[i] = i * interval;
        idx[n_all - 1] = (int)L - 1;
    }
    return n_all;
}

/* convertIntToOneLetterCode / convertIntToThreeLetterCode, src/utility.cpp:297-377, 461-: the 5-bit field can hold
 * 24..31, which fall into the switch's default branch = UNK ('X'). */
static int norm_code(int code) { return (code >= 0 && code < FCZ_NUM_CODES) ? code : FCZ_CODE_UNK; }
static int code_to_char(int code) { return FCZ_NAME1[norm_code(code)]; }
static int char_to_code(int ch) {
    for (int i = 0; i < FCZ_NUM_CODES; i++)
        if (FCZ_NAME1[i] == ch) return i;
    return FCZ_CODE_UNK;
}

/* ------------------------------------------------------------------- backbone angles before quantisation */
/* What Foldcomp::preprocess leaves in backboneTorsionAngles (getTorsionFromXYZ over the backbone, src/torsion_angle.cpp:
 * 46-96; src/foldcomp.cpp:484-485) and backboneBondAngles (Nerf::getBondAngles, src/nerf.cpp:495-508; src/foldcomp.cpp:494-
 * 495), the lists the CPython get_data(pdb_text) returns (foldcomp/foldcomp.cxx:633-671), re-cut per residue r as six
 * floats: torsions 3r, 3r+1, 3r+2 of the list (zero for the last residue), then the bond angles AT backbone atoms 3r,
 * 3r+1, 3r+2 (list index atom-1; zero at the chain's first and last atom).  Pinned to the reference's get_data() by
 * tests/test_oracle.py (oracle/_ref/pyref) and tests/golden/getdata_golden.npz. */
int fcz_oracle_backbone_angles(const uint8_t* res_type, uint32_t L, const float* xyz, float* out) {
    const f3* at = (const f3*)xyz;
    if (L < 1) return FCZ_E_LIMIT;
    f3* bb = (f3*)malloc(sizeof(f3) * 3 * L);
    uint64_t a = 0;
    for (uint32_t r = 0; r < L; r++) {
        int c = res_type[r];
        if (c >= FCZ_NUM_CODES || FCZ_NATOMS[c] == 0) { free(bb); return FCZ_E_RESIDUE; }
        for (int k = 0; k < 3; k++) bb[3 * r + k] = at[a + k];
        a += FCZ_NATOMS[c];
    }
    const uint32_t n = 3 * L;
    for (uint32_t e = 0; e < 6 * L; e++) out[e] = 0.0f;
    for (uint32_t i = 0; i + 3 < n; i++) out[6 * (i / 3) + i % 3] = dihedral(bb[i], bb[i + 1], bb[i + 2], bb[i + 3]);
    for (uint32_t m = 1; m + 1 < n; m++) out[6 * (m / 3) + 3 + m % 3] = angle3(bb[m - 1], bb[m], bb[m + 1]);
    free(bb);
    return 0;
}

/* ---------------------------------------------------------------------------------------- encode */

int64_t fcz_oracle_encode_chain(const uint8_t* res_type, uint32_t L, const float* xyz,
                                const float* bfactor, const fcz_chain_meta* meta, const char* title,
                                uint32_t title_len, int32_t b, uint8_t* out, uint64_t cap) {
    if (L < 2 || L > 65535 || b < 1) return FCZ_E_LIMIT;
    uint64_t n_atoms = 0, n_sc = 0;
    for (uint32_t r = 0; r < L; r++) {
        int c = res_type[r];
        if (c >= FCZ_NUM_CODES || FCZ_NATOMS[c] == 0) return FCZ_E_RESIDUE;
        n_atoms += FCZ_NATOMS[c];
        n_sc += FCZ_NATOMS[c] - 3;
    }
    int n_anchor = anchors(L, b, NULL);
    if (n_anchor > 255) return FCZ_E_LIMIT;
    /* src/foldcomp.cpp:1190-1214 getSize */
    uint64_t size = FCZ_HDR + 4ull * n_anchor + title_len + 36ull * n_anchor + 13 + 8ull * L + n_sc + 8 + L;
    if (!out) return (int64_t)size;
    if (size > cap) return FCZ_E_CAPACITY;

    const f3* at = (const f3*)xyz;
    uint32_t* aoff = (uint32_t*)malloc(sizeof(uint32_t) * (L + 1));
    f3* bb = (f3*)malloc(sizeof(f3) * 3 * L);
    float* ang[6]; /* phi, psi, omega, n_ca_c, ca_c_n, c_n_ca -- header order (foldcomp.cpp:1354-1365) */
    for (int k = 0; k < 6; k++) ang[k] = (float*)malloc(sizeof(float) * L);
    aoff[0] = 0;
    for (uint32_t r = 0; r < L; r++) aoff[r + 1] = aoff[r] + FCZ_NATOMS[res_type[r]];
    /* src/atom_coordinate.cpp:135-143 filterBackbone: N, CA, C = slots 0..2 */
    for (uint32_t r = 0; r < L; r++)
        for (int k = 0; k < 3; k++) bb[3 * r + k] = at[aoff[r] + k];

    /* src/foldcomp.cpp:483-505: torsions (psi,omega,phi triples) and bond angles split by i%3 */
    for (uint32_t i = 0; i + 1 < L; i++) {
        ang[1][i] = dihedral(bb[3 * i + 0], bb[3 * i + 1], bb[3 * i + 2], bb[3 * i + 3]); /* psi   */
        ang[2][i] = dihedral(bb[3 * i + 1], bb[3 * i + 2], bb[3 * i + 3], bb[3 * i + 4]); /* omega */
        ang[0][i] = dihedral(bb[3 * i + 2], bb[3 * i + 3], bb[3 * i + 4], bb[3 * i + 5]); /* phi   */
        ang[4][i] = angle3(bb[3 * i + 1], bb[3 * i + 2], bb[3 * i + 3]); /* CA-C-N  at atom 3i+2 */
        ang[5][i] = angle3(bb[3 * i + 2], bb[3 * i + 3], bb[3 * i + 4]); /* C-N-CA  at atom 3i+3 */
        ang[3][i] = angle3(bb[3 * i + 3], bb[3 * i + 4], bb[3 * i + 5]); /* N-CA-C  at atom 3i+4 */
    }
    /* src/foldcomp.cpp:508-519, src/foldcomp.h:44-49: bins */
    static const unsigned NB[6] = {4095, 4095, 2047, 255, 255, 255};
    disc_t d[6];
    for (int k = 0; k < 6; k++) d[k] = disc_fit(ang[k], L - 1, NB[k]);

    uint8_t* p = out;
    memcpy(p, "FCMP", 4);
    /* src/foldcomp.cpp:1340-1367 get_header; padding bytes (file offsets 14,15,22,23) zeroed */
    memset(p + 4, 0, 72);
    put16(p + 4, (uint16_t)L);
    put16(p + 6, meta->n_atom);
    put16(p + 8, meta->idx_residue);
    put16(p + 10, meta->idx_atom);
    p[12] = (uint8_t)n_anchor;
    p[13] = meta->chain;
    put32(p + 16, (uint32_t)n_sc);
    p[20] = (uint8_t)code_to_char(res_type[0]);     /* firstResidue (foldcomp.cpp:467) */
    p[21] = (uint8_t)code_to_char(res_type[L - 1]); /* lastResidue  (foldcomp.cpp:468) */
    put32(p + 24, title_len);
    for (int k = 0; k < 6; k++) {
        putf(p + 28 + 4 * k, d[k].min);
        putf(p + 52 + 4 * k, d[k].cont_f);
    }
    p += FCZ_HDR;
    /* src/foldcomp.cpp:1044-1049 anchor indices, title; 1051-1059 anchor atoms */
    int32_t* aidx = (int32_t*)malloc(sizeof(int32_t) * n_anchor);
    anchors(L, b, aidx);
    for (int i = 0; i < n_anchor; i++) put32(p + 4 * i, (uint32_t)aidx[i]);
    p += 4 * n_anchor;
    memcpy(p, title, title_len);
    p += title_len;
    for (int i = 0; i < n_anchor; i++)
        for (int k = 0; k < 3; k++) {
            f3 v = bb[3 * aidx[i] + k];
            putf(p, v.x); putf(p + 4, v.y); putf(p + 8, v.z);
            p += 12;
        }
    /* src/foldcomp.cpp:1061-1064 OXT */
    *p++ = meta->has_oxt;
    putf(p, meta->oxt[0]); putf(p + 4, meta->oxt[1]); putf(p + 8, meta->oxt[2]);
    p += 12;
    /* src/foldcomp.cpp:581-602 BackboneChain records; 33-52 convertBackboneChainToBytes */
    for (uint32_t i = 0; i < L; i++) {
        unsigned res = res_type[i], phi = 0, psi = 0, omg = 0, nca = 0, cac = 0, cnc = 0;
        if (i + 1 < L) {
            phi = disc_round(&d[0], ang[0][i]);
            psi = disc_round(&d[1], ang[1][i]);
            omg = disc_round(&d[2], ang[2][i]);
            nca = disc_round(&d[3], ang[3][i]);
            cac = disc_round(&d[4], ang[4][i]);
            cnc = disc_round(&d[5], ang[5][i]);
        }
        /* bit-field truncation of struct BackboneChain (src/foldcomp.h:71-81) */
        res &= 0x1F; omg &= 0x7FF; psi &= 0xFFF; phi &= 0xFFF; cac &= 0xFF; cnc &= 0xFF; nca &= 0xFF;
        p[0] = (uint8_t)((res << 3) | (omg >> 8));
        p[1] = (uint8_t)(omg & 0xFF);
        p[2] = (uint8_t)(psi >> 4);
        p[3] = (uint8_t)(((psi & 0xF) << 4) | (phi >> 8));
        p[4] = (uint8_t)(phi & 0xFF);
        p[5] = (uint8_t)cac;
        p[6] = (uint8_t)cnc;
        p[7] = (uint8_t)nca;
        p += 8;
    }
    /* src/sidechain.cpp:149-180 + src/foldcomp.cpp:532-538: one truncated byte per non-backbone atom */
    const float sc_min = (float)-180.0;
    const float sc_disc_f = (float)255u / ((float)180.0 - sc_min); /* discretizer.h:91-97 */
    for (uint32_t r = 0; r < L; r++) {
        int c = res_type[r];
        const f3* ra = at + aoff[r];
        for (int k = 3; k < FCZ_NATOMS[c]; k++) {
            unsigned pr = FCZ_PRED[c][k];
            float t = dihedral(ra[pr & 15], ra[(pr >> 4) & 15], ra[(pr >> 8) & 15], ra[k]);
            *p++ = (uint8_t)disc_trunc(sc_min, sc_disc_f, t);
        }
    }
    /* src/foldcomp.cpp:543-550, 1098-1107: B-factor discretiser + bytes */
    disc_t db = disc_fit(bfactor, L, 255);
    putf(p, db.min); putf(p + 4, db.cont_f);
    p += 8;
    for (uint32_t r = 0; r < L; r++) *p++ = (uint8_t)disc_round(&db, bfactor[r]);

    free(aidx);
    for (int k = 0; k < 6; k++) free(ang[k]);
    free(bb); free(aoff);
    return (int64_t)(p - out);
}

/* ---------------------------------------------------------------------------------------- decode */

typedef struct {
    uint32_t L, n_sc, title_len;
    int n_anchor;
    const uint8_t *aidx, *title, *anchor_xyz, *oxt, *records, *sc, *temp;
} view_t;

/* src/foldcomp.cpp:904-1036 read(): section offsets only */
static int parse(const uint8_t* b, uint64_t len, view_t* v) {
    if (len < FCZ_HDR || memcmp(b, "FCMP", 4) != 0) return FCZ_E_MAGIC;
    v->L = get16(b + 4);
    v->n_anchor = b[12];
    v->n_sc = get32(b + 16);
    v->title_len = get32(b + 24);
    uint64_t o = FCZ_HDR;
    v->aidx = b + o; o += 4ull * v->n_anchor;
    v->title = b + o; o += v->title_len;
    v->anchor_xyz = b + o; o += 36ull * v->n_anchor;
    v->oxt = b + o; o += 13;
    v->records = b + o; o += 8ull * v->L;
    v->sc = b + o; o += v->n_sc;
    v->temp = b + o; o += 8ull + v->L;
    if (o > len || v->L < 2 || v->n_anchor < 2) return FCZ_E_TRUNCATED;
    return FCZ_OK;
}

int fcz_oracle_peek(const uint8_t* blob, uint64_t len, uint32_t* L, uint64_t* n_atoms, uint32_t* title_len) {
    view_t v;
    int rc = parse(blob, len, &v);
    if (rc) return rc;
    uint64_t na = 0, nsc = 0;
    for (uint32_t r = 0; r < v.L; r++) {
        int c = norm_code(v.records[8 * r] >> 3);
        if (FCZ_NATOMS[c] == 0) return FCZ_E_RESIDUE;
        na += FCZ_NATOMS[c];
        nsc += FCZ_NATOMS[c] - 3;
    }
    if (nsc != v.n_sc) return FCZ_E_TRUNCATED; /* checkValidity E_SIDECHAIN_COUNT_MISMATCH (foldcomp.cpp:1492-1561) */
    *L = v.L; *n_atoms = na; *title_len = v.title_len;
    return FCZ_OK;
}

typedef struct { float phi, psi, omega, n_ca_c, ca_c_n, c_n_ca; int code; } rec_t;

static f3 getf3(const uint8_t* p) { f3 v = {getf(p), getf(p + 4), getf(p + 8)}; return v; }

int fcz_oracle_decode_chain(const uint8_t* blob, uint64_t len, int use_alt, uint8_t* res_type,
                            float* bfactor, float* xyz, fcz_chain_meta* meta, char* title) {
    view_t v;
    int rc = parse(blob, len, &v);
    if (rc) return rc;
    const uint32_t L = v.L;
    const int n_anchor = v.n_anchor;
    float mins[6], cfs[6];
    for (int k = 0; k < 6; k++) { mins[k] = getf(blob + 28 + 4 * k); cfs[k] = getf(blob + 52 + 4 * k); }

    /* src/foldcomp.cpp:60-77 convertBytesToBackboneChain + 122-153 decompressBackboneChain */
    rec_t* rec = (rec_t*)malloc(sizeof(rec_t) * L);
    for (uint32_t i = 0; i < L; i++) {
        const uint8_t* b = v.records + 8 * i;
        unsigned res = (unsigned)norm_code(b[0] >> 3);
        unsigned omg = ((b[0] & 7u) << 8) | b[1];
        unsigned psi = ((unsigned)b[2] << 4) | (b[3] >> 4);
        unsigned phi = ((b[3] & 0xFu) << 8) | b[4];
        rec[i].code = (int)res;
        rec[i].phi = cont(phi, mins[0], cfs[0]);
        rec[i].psi = cont(psi, mins[1], cfs[1]);
        rec[i].omega = cont(omg, mins[2], cfs[2]);
        rec[i].n_ca_c = cont(b[7], mins[3], cfs[3]);
        rec[i].ca_c_n = cont(b[5], mins[4], cfs[4]);
        rec[i].c_n_ca = cont(b[6], mins[5], cfs[5]);
        res_type[i] = (uint8_t)res;
        if (FCZ_NATOMS[res] == 0) { free(rec); return FCZ_E_RESIDUE; }
    }
    /* src/foldcomp.cpp:788-793: torsion list psi,omega,phi for records 0..L-2 */
    const uint32_t nT = 3 * (L - 1);
    float* tors = (float*)malloc(sizeof(float) * (nT ? nT : 1));
    for (uint32_t i = 0; i + 1 < L; i++) {
        tors[3 * i] = rec[i].psi; tors[3 * i + 1] = rec[i].omega; tors[3 * i + 2] = rec[i].phi;
    }
    f3* bb = (f3*)malloc(sizeof(f3) * 3 * L);    /* final backbone */
    f3* seg = (f3*)malloc(sizeof(f3) * 3 * (L + 1));
    f3* rev = (f3*)malloc(sizeof(f3) * 3 * (L + 1));
    float* bang = (float*)malloc(sizeof(float) * 3 * (L + 1));
    uint32_t n_out = 0;
    f3 prev[3];
    for (int k = 0; k < 3; k++) prev[k] = getf3(v.anchor_xyz + 12 * k); /* anchor 0 = prevAtoms */

    /* src/foldcomp.cpp:812-858 segment loop */
    for (int s = 0; s < n_anchor - 1; s++) {
        int max_index = (int)L - 1;
        int a0 = (int)get32(v.aidx + 4 * s), a1 = (int)get32(v.aidx + 4 * (s + 1));
        int first = a0 < max_index ? a0 : max_index;
        int last = (a1 + 1) < max_index ? (a1 + 1) : max_index;
        int total = last - first;              /* records [first,last) */
        if (s == n_anchor - 2) total += 1;     /* + the last record (foldcomp.cpp:828-830) */
        if (total < 1) total = 1;
        /* src/foldcomp.cpp:167-246 reconstructBackboneAtoms */
        seg[0] = prev[0]; seg[1] = prev[1]; seg[2] = prev[2];
        for (int i = 0; i < total - 1; i++) {
            const rec_t* r = &rec[first + i];
            f3 p0 = seg[3 * i], p1 = seg[3 * i + 1], p2 = seg[3 * i + 2];
            f3 n = place_atom(p0, p1, p2, (float)1.3311, r->ca_c_n, r->psi);
            float n_ca = (code_to_char(r->code) != 'P') ? (float)1.4581 : (float)1.353; /* foldcomp.cpp:204-212 */
            f3 ca = place_atom(p1, p2, n, n_ca, r->c_n_ca, r->omega);
            f3 c = place_atom(p2, n, ca, (float)1.5281, r->n_ca_c, r->phi);
            seg[3 * i + 3] = n; seg[3 * i + 4] = ca; seg[3 * i + 5] = c;
        }
        int n = 3 * total;
        /* torsion subset (foldcomp.cpp:833-844) */
        int tmax = (int)nT - 1;
        int tf = a0 * 3 < tmax ? a0 * 3 : tmax;
        int tl = a1 * 3 < tmax ? a1 * 3 : tmax;
        int nt = tl - tf;
        if (s == n_anchor - 2) nt += 1;
        /* src/foldcomp.cpp:248-273 reconstructBackboneReverse */
        for (int i = 0; i < n; i++) rev[i] = seg[i];
        for (int k = 0; k < 3; k++) rev[n - 3 + k] = getf3(v.anchor_xyz + 36 * (s + 1) + 12 * k);
        /* src/nerf.cpp:495-508 getBondAngles of the FORWARD atoms: bang[j] = angle at atom j+1 */
        for (int i = 1; i < n - 1; i++) bang[i - 1] = angle3(seg[i - 1], seg[i], seg[i + 1]);
        int nb = n - 2;
        /* src/nerf.cpp:342-379 reconstructWithReversed, written with forward indices:
         * reversed atom R[i] = rev[n-1-i]; reversed torsion rt[i] = tors[tf + nt-1-i];
         * reversed bond angle rb[i] = bang[nb-1-i]. */
        for (int i = 0; i < n - 3; i++) {
            int q = n - 4 - i; /* forward index of the atom being placed */
            /* bond length by names curr_TO_prev (nerf.cpp:363-364, nerf.h:37-43); atom kind = q%3 */
            float bl = (q % 3 == 0) ? (float)1.4581 /* N_TO_CA */
                     : (q % 3 == 1) ? (float)1.5281 /* CA_TO_C */
                                    : (float)1.3311 /* C_TO_N  */;
            float ba = bang[nb - 1 - (i + 1)];
            float ta = tors[tf + nt - 1 - i];
            rev[q] = place_atom(rev[q + 3], rev[q + 2], rev[q + 1], bl, ba, ta);
        }
        /* src/atom_coordinate.cpp:145-163 weightedAverage */
        for (int i = 0; i < n; i++) {
            seg[i].x = ((seg[i].x * (float)(n - i)) + (rev[i].x * (float)i)) / (float)n;
            seg[i].y = ((seg[i].y * (float)(n - i)) + (rev[i].y * (float)i)) / (float)n;
            seg[i].z = ((seg[i].z * (float)(n - i)) + (rev[i].z * (float)i)) / (float)n;
        }
        /* src/foldcomp.cpp:848-857 append all but the last 3 (all for the last segment); next start */
        int keep = (s != n_anchor - 2) ? n - 3 : n;
        for (int i = 0; i < keep && n_out < 3 * L; i++) bb[n_out++] = seg[i];
        prev[0] = seg[n - 3]; prev[1] = seg[n - 2]; prev[2] = seg[n - 1];
    }

    /* src/foldcomp.cpp:860-880 side chains: src/nerf.cpp:106-155 reconstructAminoAcid;
     * torsion = FixedAngleDiscretizer(255).continuize(byte) (foldcomp.cpp:338-369) */
    const float sc_min = (float)-180.0;
    const float sc_cont_f = ((float)180.0 - sc_min) / (float)255u;
    disc_t db;
    db.min = getf(v.temp); db.cont_f = getf(v.temp + 4);
    f3* out = (f3*)xyz;
    uint64_t a = 0, t = 0;
    for (uint32_t r = 0; r < L; r++) {
        int c = rec[r].code;
        int na = FCZ_NATOMS[c];
        f3 ra[FCZ_MAX_ATOMS];
        ra[0] = bb[3 * r]; ra[1] = bb[3 * r + 1]; ra[2] = bb[3 * r + 2];
        for (int k = 3; k < na; k++) {
            unsigned pr = FCZ_PRED[c][k];
            float tor = cont(v.sc[t++], sc_min, sc_cont_f);
            ra[k] = place_atom(ra[pr & 15], ra[(pr >> 4) & 15], ra[(pr >> 8) & 15], FCZ_BLEN[c][k],
                               FCZ_BANG[c][k], tor);
        }
        for (int k = 0; k < na; k++) out[a + k] = use_alt ? ra[FCZ_ALT[c][k]] : ra[k]; /* foldcomp.cpp:1563-1577 */
        a += na;
        bfactor[r] = cont(v.temp[8 + r], db.min, db.cont_f); /* foldcomp.cpp:884-886 */
    }
    if (meta) {
        meta->n_atom = get16(blob + 6);
        meta->idx_residue = get16(blob + 8);
        meta->idx_atom = get16(blob + 10);
        meta->chain = blob[13];
        meta->has_oxt = v.oxt[0];
        meta->oxt[0] = getf(v.oxt + 1); meta->oxt[1] = getf(v.oxt + 5); meta->oxt[2] = getf(v.oxt + 9);
    }
    if (title) memcpy(title, v.title, v.title_len);
    free(bang); free(rev); free(seg); free(bb); free(tors); free(rec);
    (void)char_to_code;
    return FCZ_OK;
}

/* ---------------------------------------------------------------------------------------- batches */

int fcz_oracle_encode_batch(const fcz_chain_batch* in, fcz_blob_batch* out, int32_t b, int n_threads) {
    uint32_t n = in->n_chains;
    out->n_chains = n;
    out->blob_off[0] = 0;
    for (uint32_t c = 0; c < n; c++) {
        uint32_t L = in->res_off[c + 1] - in->res_off[c];
        int64_t sz = fcz_oracle_encode_chain(in->res_type + in->res_off[c], L, NULL, NULL, NULL, NULL,
                                             in->title_off[c + 1] - in->title_off[c], b, NULL, 0);
        if (out->status) out->status[c] = sz < 0 ? (int32_t)sz : FCZ_OK;
        out->blob_off[c + 1] = out->blob_off[c] + (sz < 0 ? 0 : (uint64_t)sz);
    }
    if (out->blob_off[n] > out->bytes_cap) return FCZ_E_CAPACITY;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (uint32_t c = 0; c < n; c++) {
        uint64_t sz = out->blob_off[c + 1] - out->blob_off[c];
        if (!sz) continue;
        uint32_t L = in->res_off[c + 1] - in->res_off[c];
        fcz_oracle_encode_chain(in->res_type + in->res_off[c], L, in->xyz + 3 * in->atom_off[c],
                                in->bfactor + in->res_off[c], &in->meta[c], in->titles + in->title_off[c],
                                in->title_off[c + 1] - in->title_off[c], b, out->bytes + out->blob_off[c], sz);
    }
    return FCZ_OK;
}

int fcz_oracle_decode_plan(const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals) {
    uint32_t n = in->n_chains;
    out->n_chains = n;
    out->res_off[0] = 0; out->atom_off[0] = 0; out->title_off[0] = 0;
    for (uint32_t c = 0; c < n; c++) {
        uint32_t L = 0, tl = 0;
        uint64_t na = 0;
        int rc = fcz_oracle_peek(in->bytes + in->blob_off[c], in->blob_off[c + 1] - in->blob_off[c], &L, &na, &tl);
        if (rc) { L = 0; na = 0; tl = 0; }
        if (out->status) out->status[c] = rc;
        out->res_off[c + 1] = out->res_off[c] + L;
        out->atom_off[c + 1] = out->atom_off[c] + na;
        out->title_off[c + 1] = out->title_off[c] + tl;
    }
    totals->n_res = out->res_off[n];
    totals->n_atoms = out->atom_off[n];
    totals->n_title_bytes = out->title_off[n];
    totals->n_blob_bytes = in->blob_off[n];
    return FCZ_OK;
}

int fcz_oracle_decode_batch(const fcz_blob_batch* in, fcz_chain_batch* out, int use_alt, int n_threads) {
    uint32_t n = in->n_chains;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (uint32_t c = 0; c < n; c++) {
        if (out->res_off[c + 1] == out->res_off[c]) continue;
        int rc = fcz_oracle_decode_chain(in->bytes + in->blob_off[c], in->blob_off[c + 1] - in->blob_off[c], use_alt,
                                         out->res_type + out->res_off[c], out->bfactor + out->res_off[c],
                                         out->xyz + 3 * out->atom_off[c], &out->meta[c],
                                         out->titles ? out->titles + out->title_off[c] : NULL);
        if (out->status) out->status[c] = rc;
    }
    return FCZ_OK;
}

/* ------------------------------------------------------------------------------ text (section 8 f1/f4) */

#include <stdio.h>

/* src/atom_coordinate.cpp:173-183 itoa_pos_only */
static void itoa_pos_only(int n, char* s) {
    int i = 0;
    do { s[i++] = (char)(n % 10 + '0'); } while ((n /= 10) > 0);
    s[i] = 0;
    for (int a = 0, b = i - 1; a < b; a++, b--) { char c = s[a]; s[a] = s[b]; s[b] = c; }
}

/* src/atom_coordinate.cpp:186-218 fast_ftoa<T,P>: float arithmetic throughout */
static void fast_ftoa(float n, int T, int P, char* s) {
    float rounded = n + ((n < 0) ? -(0.5f / (float)T) : (0.5f / (float)T));
    int32_t integer = (int32_t)rounded;
    int32_t decimal = (int32_t)((rounded - (float)integer) * (float)T);
    char* data = s;
    if (n < 0) {
        integer = abs(integer);
        decimal = abs(decimal);
        *data++ = '-';
    }
    itoa_pos_only(integer, data);
    data += strlen(data);
    *data++ = '.';
    char buffer[16];
    itoa_pos_only(decimal, buffer);
    int len = (int)strlen(buffer);
    for (int i = 0; i < P - len; i++) *data++ = '0';
    memcpy(data, buffer, (size_t)len);
    data[len] = 0;
}

typedef struct { char* p; uint64_t cap, n; } sink_t;
static void emit(sink_t* o, const char* s, size_t len) {
    for (size_t i = 0; i < len; i++, o->n++)
        if (o->n < o->cap) o->p[o->n] = s[i];
}
static void emitf(sink_t* o, const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n > 0) emit(o, buf, (size_t)(n < (int)sizeof buf ? n : (int)sizeof buf - 1));
}

/* src/atom_coordinate.cpp:246-275: one ATOM line; std::setw is a minimum width */
static void atom_line(sink_t* o, int serial, const char* atom, const char* res3, char chain, int resnum,
                      float x, float y, float z, float b) {
    char fx[32], fy[32], fz[32], fb[32];
    fast_ftoa(x, 1000, 3, fx);
    fast_ftoa(y, 1000, 3, fy);
    fast_ftoa(z, 1000, 3, fz);
    fast_ftoa(b, 100, 2, fb);
    if (strlen(atom) == 4) emitf(o, "ATOM  %5d %-4s %3s %c%4d    %8s%8s%8s  1.00%6s          %2c  \n", serial, atom, res3, chain, resnum, fx, fy, fz, fb, atom[0]);
    else emitf(o, "ATOM  %5d  %-3s %3s %c%4d    %8s%8s%8s  1.00%6s          %2c  \n", serial, atom, res3, chain, resnum, fx, fy, fz, fb, atom[0]);
}

/* writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291) over a chain in the decoder's output layout,
 * labelled as Foldcomp::decompress labels its atoms (serials from idxAtom, src/atom_coordinate.cpp:356-360;
 * residue_index idxResidue + r; the OXT record with residue_index = nResidue, src/foldcomp.cpp:958-961).
 * Returns the text length; at most cap bytes are written. */
int64_t fcz_oracle_format_pdb(const uint8_t* res_type, uint32_t L, const float* xyz, const float* bfactor,
                              const fcz_chain_meta* meta, const char* title, uint32_t title_len, int use_alt,
                              char* out, uint64_t cap) {
    sink_t o = {out, cap, 0};
    if (title_len) { /* src/atom_coordinate.cpp:223-243 */
        int remaining = (int)title_len;
        emitf(&o, "TITLE     %.*s\n", remaining < 70 ? remaining : 70, title);
        remaining -= 70;
        int continuation = 2;
        while (remaining > 0) {
            emitf(&o, "TITLE  % 3d%.*s\n", continuation, remaining < 70 ? remaining : 70, title + ((int)title_len - remaining));
            remaining -= 70;
            continuation++;
        }
    }
    int serial = meta->idx_atom;
    uint64_t a = 0;
    int last_serial = 0, last_resnum = 0;
    const char* last_res3 = "";
    int any = 0;
    for (uint32_t r = 0; r < L; r++) {
        const int code = res_type[r];
        for (int k = 0; k < FCZ_NATOMS[code]; k++, a++) {
            const char* name = FCZ_ATOM_NAME[code][use_alt ? FCZ_ALT[code][k] : k];
            atom_line(&o, serial, name, FCZ_NAME3[code], (char)meta->chain, (int)meta->idx_residue + (int)r,
                      xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2], bfactor[r]);
            last_serial = serial; last_resnum = (int)meta->idx_residue + (int)r; last_res3 = FCZ_NAME3[code]; any = 1;
            serial++;
        }
    }
    if (meta->has_oxt && L) {
        const int code = res_type[L - 1];
        atom_line(&o, serial, "OXT", FCZ_NAME3[code], (char)meta->chain, (int)L, meta->oxt[0], meta->oxt[1], meta->oxt[2], bfactor[L - 1]);
        last_serial = serial; last_resnum = (int)L; last_res3 = FCZ_NAME3[code]; any = 1;
    }
    if (any) emitf(&o, "TER   %5d      %3s %c%4d\n", last_serial + 1, last_res3, (char)meta->chain, last_resnum);
    return (int64_t)o.n;
}

/* Foldcomp::extract (src/foldcomp.cpp:1260-1336) */
int64_t fcz_oracle_extract(const uint8_t* blob, uint64_t len, int type, int digits, char* out, uint64_t cap) {
    view_t v;
    int rc = parse(blob, len, &v);
    if (rc) return rc;
    sink_t o = {out, cap, 0};
    if (type == 1) {
        for (uint32_t r = 0; r < v.L; r++) { char ch = (char)code_to_char(v.records[8 * r] >> 3); emit(&o, &ch, 1); }
        return (int64_t)o.n;
    }
    if (digits < 1) digits = 1; else if (digits > 4) digits = 4;
    const float tmin = getf(v.temp), tcont = getf(v.temp + 4);
    const float maxval = (float)((double)tcont * (pow(2, 8) - 1) + (double)tmin);
    const int zero_to_one = (maxval <= 1.0f && digits <= 2);
    for (uint32_t i = 0; i < v.L; i++) {
        const float t = cont(v.temp[8 + i], tmin, tcont);
        float clamped;
        char d1, d2;
        if (zero_to_one) {
            clamped = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
            d1 = (char)((int)(clamped * 10.0f) % 10) + '0';
            d2 = (char)((int)(clamped * 100.0f) % 10) + '0';
        } else {
            clamped = t < 0.0f ? 0.0f : (t > 100.0f ? 100.0f : t);
            d1 = (char)(clamped / 10.0f) + '0';
            d2 = (char)((int)clamped % 10) + '0';
        }
        emit(&o, &d1, 1);
        if (digits > 1) emit(&o, &d2, 1);
        if (digits >= 3) {
            char d3 = (char)((int)(clamped * 10.0f) % 10) + '0';
            emit(&o, ".", 1);
            emit(&o, &d3, 1);
        }
        if (digits == 4) {
            char d4 = (char)((int)(clamped * 100.0f) % 10) + '0';
            emit(&o, &d4, 1);
        }
        if (digits > 1 && i != v.L - 1) emit(&o, ",", 1);
    }
    return (int64_t)o.n;
}

/* Continuised backbone angles per residue record: decompressBackboneChain (src/foldcomp.cpp:122-153) with _continuize
 * (155-158).  out gets 6 floats per residue: phi, psi, omega, N-CA-C, CA-C-N, C-N-CA.  Returns the residue count. */
int64_t fcz_oracle_unpack_angles(const uint8_t* blob, uint64_t len, float* out) {
    view_t v;
    int rc = parse(blob, len, &v);
    if (rc) return rc;
    float mins[6], cfs[6];
    for (int k = 0; k < 6; k++) { mins[k] = getf(blob + 28 + 4 * k); cfs[k] = getf(blob + 52 + 4 * k); }
    for (uint32_t r = 0; r < v.L; r++) {
        const uint8_t* b = v.records + 8 * r; /* convertBytesToBackboneChain, src/foldcomp.cpp:60-77 */
        unsigned omega = ((b[0] & 7u) << 8) | b[1];
        unsigned psi = ((unsigned)b[2] << 4) | (b[3] >> 4);
        unsigned phi = ((b[3] & 15u) << 8) | b[4];
        out[6 * r + 0] = cont(phi, mins[0], cfs[0]);
        out[6 * r + 1] = cont(psi, mins[1], cfs[1]);
        out[6 * r + 2] = cont(omega, mins[2], cfs[2]);
        out[6 * r + 3] = cont(b[7], mins[3], cfs[3]);
        out[6 * r + 4] = cont(b[5], mins[4], cfs[4]);
        out[6 * r + 5] = cont(b[6], mins[5], cfs[5]);
    }
    return (int64_t)v.L;
}

/* Foldcomp::read (src/foldcomp.cpp:904-1036) + Foldcomp::checkValidity (src/foldcomp.cpp:1492-1532).  read() sizes its
 * three vectors from the header (compressedBackBone.resize(nResidue) 979-984, nSideChainTorsion push_backs 1010-1014,
 * nResidue push_backs 1027-1031), so the three COUNT_MISMATCH classes cannot occur on a blob that is complete; they
 * are used here for a blob that ends before the section (the reference reads past the end of its stream there and
 * checks whatever its buffers held).  Then, in this order: every record has phi = psi = omega = 0 (1499-1505), every
 * side-chain byte is 0 (1506-1510; std::all_of is true on an empty range), every B-factor byte is 0 (1511-1515).
 * Returns the read status (0, FCZ_E_MAGIC, FCZ_E_TRUNCATED); *validity receives the class 0..6. */
int fcz_oracle_check(const uint8_t* b, uint64_t len, int* validity) {
    *validity = 0;
    if (len < 4 || memcmp(b, "FCMP", 4) != 0) return FCZ_E_MAGIC;
    if (len < FCZ_HDR) { *validity = 1; return FCZ_E_TRUNCATED; }
    const uint32_t L = get16(b + 4), n_anchor = b[12], n_sc = get32(b + 16), title_len = get32(b + 24);
    uint64_t o = FCZ_HDR + 4ull * n_anchor + title_len + 36ull * n_anchor + 13;
    const uint8_t* rec = b + o;
    o += 8ull * L;
    if (o > len) { *validity = 1; return FCZ_E_TRUNCATED; }
    const uint8_t* sc = b + o;
    o += n_sc;
    if (o > len) { *validity = 2; return FCZ_E_TRUNCATED; }
    const uint8_t* temp = b + o + 8;
    o += 8ull + L;
    if (o > len) { *validity = 3; return FCZ_E_TRUNCATED; }
    int empty_bb = 1, empty_sc = 1, empty_t = 1;
    for (uint32_t r = 0; r < L; r++) {
        const uint8_t* p = rec + 8 * r;
        unsigned omg = ((p[0] & 7u) << 8) | p[1], psi = ((unsigned)p[2] << 4) | (p[3] >> 4), phi = ((p[3] & 0xFu) << 8) | p[4];
        if (phi || psi || omg) empty_bb = 0;
    }
    for (uint32_t i = 0; i < n_sc; i++) if (sc[i]) empty_sc = 0;
    for (uint32_t i = 0; i < L; i++) if (temp[i]) empty_t = 0;
    *validity = empty_bb ? 4 : (empty_sc ? 5 : (empty_t ? 6 : 0));
    return FCZ_OK;
}
