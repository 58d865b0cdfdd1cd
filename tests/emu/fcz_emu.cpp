// tests/emu/fcz_emu.cpp -- TEST HARNESS (not product, not a fallback).
// Instantiates the per-chain codec of foldcomp_b200/csrc/fcz_codec.h with a ONE-THREAD execution
// context so the algorithm the CUDA kernels run (phases, indexing, stitch decomposition, bit
// packing) can be checked against the oracle on a machine without a GPU.  It shares no code with
// oracle/; what it cannot check (barriers, TMA staging, device libm) is covered by the -m gpu tests.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../foldcomp_b200/csrc/fcz_codec.h"
#include "../../foldcomp_b200/csrc/fcz_text.h"
#include "../../foldcomp_b200/csrc/fcz_parse.h"

using namespace fcz;

struct HostCtx {
    int tid = 0, nthr = 1, lane = 0, warp = 0, nwarps = 1, wsize = 1;
    void sync() {}
    void wsync() {}
    void mark(int) {}
    void stage_wait() {}
    uint32_t excl_scan(uint32_t) { return 0; }
    float wmin(float v) { return v; }
    float wmax(float v) { return v; }
    uint32_t atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
    int32_t wmin_i(int32_t v) { return v; }
    int32_t wmax_i(int32_t v) { return v; }
    void copy_out_same_phase(char* dst, const char* src, uint32_t bytes) { copy_same_phase(*this, dst, src, bytes); }
    void atomic_min_u(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
    void atomic_or_u(uint32_t* p, uint32_t v) { *p |= v; }
    void atomic_min_i(int32_t* p, int32_t v) { if (v < *p) *p = v; }
    void atomic_max_i(int32_t* p, int32_t v) { if (v > *p) *p = v; }
};

static const Tables* tables() {
    static Tables t;
    static bool init = false;
    if (!init) { build_tables(&t); init = true; }
    return &t;
}

static thread_local uint32_t g_last_list = 0, g_last_bad = 0, g_last_cap = 0, g_last_by_array[6] = {0, 0, 0, 0, 0, 0};

extern "C" {

int64_t emu_encode_chain(const uint8_t* res_type, uint32_t L, const float* xyz, const float* bfac,
                         const fcz_chain_meta* meta, const char* title, uint32_t title_len, int32_t b,
                         uint8_t* out, uint64_t cap) {
    const Tables* tb = tables();
    uint32_t A = 0;
    for (uint32_t r = 0; r < L; r++) {
        if (res_type[r] >= FCZ_NUM_CODES || tb->natoms[res_type[r]] == 0) return FCZ_E_RESIDUE;
        A += tb->natoms[res_type[r]];
    }
    if (L < 2 || L > 65535 || b < 1 || anchor_count(L, b) > 255) return FCZ_E_LIMIT;
    Layout y = make_layout(L, A - 3 * L, title_len, (uint32_t)anchor_count(L, b));
    if (!out) return y.size;
    if (y.size > cap) return FCZ_E_CAPACITY;
    std::vector<uint32_t> aoff(L + 1);
    std::vector<uint16_t> ares(A);
    std::vector<float> ang(6 * (size_t)L), red(FCZ_RED_FLOATS(1));
    std::vector<uint32_t> fl(FCZ_FL_WORDS), list(enc_list_cap(L));
    std::vector<float> xe(enc_list_cap(L));
    EncChain ch;
    ch.L = L; ch.A = A; ch.title_len = title_len; ch.b = b;
    ch.type = res_type; ch.bfac = bfac; ch.X = xyz; ch.title = title; ch.meta = meta; ch.B = out;
    ch.aoff = aoff.data(); ch.sres = ares.data(); ch.ang = ang.data(); ch.red = red.data(); ch.fl = fl.data();
    ch.list = list.data(); ch.xe = xe.data(); ch.list_cap = enc_list_cap(L); ch.tbg = tb;
    HostCtx cx;
    encode_chain(cx, tb, ch);
    g_last_list = fl[FL_N]; g_last_bad = fl[FL_BAD];
    for (int a = 0; a < 6; a++) g_last_by_array[a] = 0;
    for (uint32_t j = 0; j < fl[FL_N] && j < ch.list_cap; j++) g_last_by_array[list[j] / L]++;
    g_last_cap = ch.list_cap;
    return y.size;
}
// float-first path of the last emu_encode_chain on this thread: values re-evaluated exactly, and whether the chain
// fell back to the all-exact path (degenerate geometry or list overflow)
void emu_last_stats(uint32_t* out) {
    out[0] = g_last_list; out[1] = (g_last_bad || g_last_list > g_last_cap) ? 1u : 0u;
    for (int a = 0; a < 6; a++) out[2 + a] = g_last_by_array[a];
}

int emu_decode_chain(const uint8_t* blob, uint64_t len, int use_alt, uint8_t* res_type, float* bfac,
                     float* xyz, fcz_chain_meta* meta, char* title) {
    const Tables* tb = tables();
    if (len < HDR_BYTES || memcmp(blob, "FCMP", 4) != 0) return FCZ_E_MAGIC;
    uint32_t L = get_u16(blob + OFF_NRES);
    Layout y = make_layout(L, get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
    if (y.size > len || L < 2 || y.n_anchor < 2) return FCZ_E_TRUNCATED;
    std::vector<uint32_t> aoff(L + 1);
    std::vector<cs> tor(3 * (size_t)L), ang(3 * (size_t)L);
    std::vector<float> seg((size_t)y.n_anchor * FCZ_SEG_FLOATS);
    std::vector<float> rev(9 * (size_t)L);
    std::vector<uint8_t> segid(L);
    DecChain ch;
    ch.blob = blob; ch.y = y; ch.use_alt = use_alt;
    ch.out_xyz = xyz; ch.out_type = res_type; ch.out_bfac = bfac; ch.out_meta = meta; ch.out_title = title;
    ch.aoff = aoff.data(); ch.tor = tor.data(); ch.ang = ang.data(); ch.seg = seg.data(); ch.rev = rev.data(); ch.segid = segid.data(); ch.loc = nullptr;
    std::vector<uint16_t> order(L); std::vector<uint32_t> bins(32);
    ch.order = order.data(); ch.bins = bins.data(); ch.codes = nullptr; ch.sc = nullptr;
    HostCtx cx;
    decode_chain(cx, tb, ch);
    return FCZ_OK;
}

// --- the certified shortcuts of the encode path, exposed for exhaustive comparison with the exact formulas
void emu_sc_bytes(const float* inner, const float* p, const uint8_t* neg, uint32_t n, uint8_t* fast, uint8_t* exact) {
    const Tables* tb = tables();
    for (uint32_t i = 0; i < n; i++) {
        DotParts d; d.inner = inner[i]; d.p = p[i];
        fast[i] = sc_byte_fast(tb, d, neg[i] != 0);
        exact[i] = (uint8_t)sc_byte_of_cos(cos_exact(d), neg[i] != 0);
    }
}
void emu_thresholds(float* pos, float* neg) {
    const Tables* tb = tables();
    memcpy(pos, tb->sc_pos, sizeof tb->sc_pos);
    memcpy(neg, tb->sc_neg, sizeof tb->sc_neg);
}
uint64_t emu_cos_deg_mismatches(const float* inner, const float* p, uint32_t n, uint64_t* fallbacks) {
    uint64_t bad = 0, fb = 0;
    for (uint32_t i = 0; i < n; i++) {
        DotParts d; d.inner = inner[i]; d.p = p[i];
        const float ce = cos_exact(d), cr = cos_ref(d);
        float tmp;
        if (!(d.p >= 1e-30f && d.p <= 1e30f && same_float((double)d.inner * drsqrt_((double)d.p), &tmp))) fb++;
        if (memcmp(&ce, &cr, 4) != 0 && !(ce != ce && cr != cr)) bad++;
        const double ac = acos((double)ce);
        const float de = (float)(ac * 180.0 / M_PI), dr = deg_ref(ac);
        if (memcmp(&de, &dr, 4) != 0 && !(de != de && dr != dr)) bad++;
    }
    *fallbacks = fb;
    return bad;
}

// PDB text of one chain through the product's plan + emit units (fcz_text.h), one host thread.
int64_t emu_format_pdb(const uint8_t* res_type, uint32_t L, const float* xyz, const float* bfac, const fcz_chain_meta* meta,
                       const char* title, uint32_t title_len, int use_alt, char* out, uint64_t cap) {
    static TextTables tt;
    static bool init = false;
    if (!init) { build_text_tables(&tt); init = true; }
    uint32_t A = 0;
    for (uint32_t r = 0; r < L; r++) A += tt.natoms[res_type[r]];
    std::vector<uint32_t> aoff(L + 1), toff(L + 1);
    PdbChain ch;
    ch.L = L; ch.A = A; ch.title_len = title_len; ch.type = res_type; ch.bfac = bfac; ch.X = xyz; ch.title = title; ch.meta = meta;
    ch.use_alt = use_alt; ch.aoff = aoff.data(); ch.toff = toff.data();
    HostCtx cx;
    uint32_t scratch = 0;
    const uint32_t total = pdb_plan_chain(cx, &tt, ch, &scratch);
    if (!out || total > cap) return total;
    std::vector<V16> stage(FCZ_PDB_STAGE_BYTES / 16 + 1);
    memset(out, '#', total);  // every byte must be written by the emit units
    uint32_t us[2 * FCZ_PDB_UNIT_RES + 2];
    for (uint32_t r = 0; r < L; r += FCZ_PDB_UNIT_RES)
        pdb_emit_unit(cx, &tt, ch, r, r + FCZ_PDB_UNIT_RES < L ? r + FCZ_PDB_UNIT_RES : L, out, (char*)stage.data(), us);
    return total;
}

int64_t emu_extract(const uint8_t* blob, uint64_t len, int type, int digits, char* out, uint64_t cap) {
    static TextTables tt;
    static bool init = false;
    if (!init) { build_text_tables(&tt); init = true; }
    if (len < HDR_BYTES) return FCZ_E_MAGIC;
    const Layout y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
    if (y.size > len) return FCZ_E_TRUNCATED;
    if (digits < 1) digits = 1; else if (digits > 4) digits = 4;
    const uint32_t n = extract_len(y.L, type, (uint32_t)digits);
    if (!out || n > cap) return n;
    HostCtx cx;
    extract_chain(cx, &tt, blob, y, type, (uint32_t)digits, out);
    return n;
}

}  // extern "C"

// Exhaustive (stride = 1) or sampled check of acos_deg_fast against long double over the floats in [-1, 1]:
// returns the maximum relative error as a base-2 exponent * 1000 (e.g. -46500 = 2^-46.5) and counts the inputs
// whose certified float differs from the reference's float (must be 0) and the uncertified ones.
extern "C" long emu_acos_check(uint32_t stride, uint64_t* n_wrong, uint64_t* n_uncertified) {
    double worst = 0;
    uint64_t wrong = 0, unc = 0;
    const uint32_t one = 0x3f800000u;
#pragma omp parallel for reduction(max : worst) reduction(+ : wrong, unc) schedule(static)
    for (int64_t k = 0; k <= (int64_t)one; k += stride) {
        for (int sgn = 0; sgn < 2; sgn++) {
            uint32_t u = (uint32_t)k | (sgn ? 0x80000000u : 0u);
            float c;
            memcpy(&c, &u, 4);
            const long double ref = acosl((long double)c) * 180.0L / 3.14159265358979323846264338327950288L;
            const double v = fcz::acos_deg_fast(c);
            if (ref != 0) {
                const double rel = (double)fabsl(((long double)v - ref) / ref);
                if (rel > worst) worst = rel;
            } else if (v != 0) worst = 1;
            float f;
            const float want = (float)(acos((double)c) * 180.0 / M_PI);
            if (fcz::acos_deg_certified(c, &f)) { if (memcmp(&f, &want, 4) != 0) wrong++; } else unc++;
        }
    }
    *n_wrong = wrong;
    *n_uncertified = unc;
    return worst > 0 ? (long)(1000.0 * log2(worst)) : -99999;
}

// cossin_deg (decode) against double sin/cos: n points evenly over [lo, hi] degrees; returns the maximum
// absolute error of either component in units of 1e-9.
extern "C" long emu_cossin_check(double lo, double hi, uint64_t n) {
    double worst = 0;
#pragma omp parallel for reduction(max : worst) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float deg = (float)(lo + (hi - lo) * (double)i / (double)(n - 1));
        const fcz::cs v = fcz::cossin_deg(deg);
        const double r = (double)deg * (M_PI / 180.0);
        const double e = fmax(fabs((double)v.c - cos(r)), fabs((double)v.s - sin(r)));
        if (e > worst) worst = e;
    }
    return (long)(worst * 1e9);
}

// acosdeg_f (the encoder's float-first arccosine) against the reference's float (float)(acos((double)c) * 180.0 / M_PI)
// over the floats in [-1, 1] (stride 1 = all of them), with the reciprocal square root it is handed perturbed by
// -4 .. +4 ulp: returns the largest absolute deviation in degrees.
extern "C" double emu_acosdeg_f_check(uint32_t stride) {
    double worst = 0;
    const uint32_t one = 0x3f800000u;
#pragma omp parallel for reduction(max : worst) schedule(static)
    for (int64_t k = 0; k <= (int64_t)one; k += stride) {
        for (int sgn = 0; sgn < 2; sgn++) {
            uint32_t u = (uint32_t)k | (sgn ? 0x80000000u : 0u);
            float c;
            memcpy(&c, &u, 4);
            const float want = (float)(acos((double)c) * 180.0 / M_PI);
            const float zb = (1.0f - fabsf(c)) * 0.5f;
            const float rs0 = (float)(1.0 / sqrt((double)(zb > 1e-30f ? zb : 1e-30f)));
            for (int d = -4; d <= 4; d += 4) {
                float rs = rs0;
                for (int i = 0; i < (d < 0 ? -d : d); i++) rs = nextafterf(rs, d < 0 ? 0.0f : INFINITY);
                const double e = fabs((double)fcz::acosdeg_f(c, rs) - (double)want);
                if (e > worst) worst = e;
            }
        }
    }
    return worst;
}

// The GPU parser's per-entry algorithm (fcz_parse.h) on one host thread: returns the parser's flag (0 ok), or -1 when
// the output arrays are too small.
extern "C" int emu_parse_pdb(const char* text, uint32_t len, uint8_t* res_type, float* bfac, float* xyz, fcz_chain_meta* meta,
                             uint32_t* n_res, uint32_t* n_atoms, uint32_t cap_res, uint32_t cap_atoms) {
    static ParseTables pt;
    static bool init = false;
    if (!init) { build_parse_tables(&pt); init = true; }
    uint32_t max_lines = 1;
    for (uint32_t i = 0; i < len; i++) max_lines += text[i] == '\n';
    std::vector<uint32_t> lines(max_lines + 1), rstart(max_lines + 1), scratch(8);
    std::vector<RawAtom> raw(2 * (size_t)max_lines);
    ParseEntry e;
    e.text = text; e.len = len; e.max_lines = max_lines; e.lines = lines.data(); e.raw = raw.data(); e.rstart = rstart.data(); e.scratch = scratch.data();
    HostCtx cx;
    parse_entry_plan(cx, &pt, e);
    if (scratch[PS_FLAG]) return (int)scratch[PS_FLAG];
    *n_res = scratch[PS_NRES]; *n_atoms = scratch[PS_NSLOT];
    if (*n_res > cap_res || *n_atoms > cap_atoms) return -1;
    parse_entry_emit(cx, &pt, e, res_type, bfac, xyz, meta);
    return 0;
}
// parse_fixed_float against strtof on n fields of width w (concatenated): returns the number of mismatches; *rejected = fields
// the fast shape does not cover
extern "C" uint64_t emu_parse_float_check(const char* fields, uint32_t w, uint64_t n, uint64_t* rejected) {
    uint64_t bad = 0, rej = 0;
#pragma omp parallel for reduction(+ : bad, rej) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        char buf[40];
        memcpy(buf, fields + (size_t)i * w, w);
        buf[w] = 0;
        bool ok;
        const float got = parse_fixed_float(buf, w, &ok);
        if (!ok) { rej++; continue; }
        const float want = strtof(buf, nullptr);
        if (memcmp(&got, &want, 4) != 0) bad++;
    }
    *rejected = rej;
    return bad;
}

// enc_raw_angles (the body of k_raw_angles) on one host thread
extern "C" int emu_backbone_angles(const uint8_t* res_type, uint32_t L, const float* xyz, float* out) {
    std::vector<uint32_t> aoff(L + 1, 0);
    for (uint32_t r = 0; r < L; r++) aoff[r + 1] = aoff[r] + (uint32_t)FCZ_NATOMS[res_type[r] < FCZ_NUM_CODES ? res_type[r] : FCZ_CODE_UNK];
    EncChain ch;
    memset(&ch, 0, sizeof ch);
    ch.L = L; ch.X = xyz; ch.aoff = aoff.data(); ch.type = res_type;
    HostCtx cx;
    enc_raw_angles(cx, ch, out);
    return 0;
}
