// foldcomp_b200/csrc/fcz_parse.h -- fixed-column PDB text -> canonical slot layout, per entry, written against the
// abstract execution context of fcz_codec.h (host + device): SURVEY.md section 8 f3 on the GPU.
//
// What it replaces, for ONE single-chain PDB text:
//   foldcomp/foldcomp.cxx:253-293   the ATOM-record parser of the CPython compress(): fields by fixed columns
//                                   (atom 12-15, residue 17-19, chain 21, serial 6-10, residue number 22-25,
//                                   x y z 30-53, B-factor 60-65), stoi / stof per field, "Multiple chains" (flag 2),
//                                   "No ATOM" (flag 1)
//   src/atom_coordinate.cpp:362-370 removeAlternativePosition: an atom named like its predecessor is dropped
//   src/atom_coordinate.cpp:304-328 splitAtomByResidue: a new residue where the residue number changes, the last
//                                   atom always joins the current residue
//   src/sidechain.cpp:140-147, src/foldcomp.cpp:473-481, 543-547  the encoder's by-name lookups: first atom of each
//                                   table name (missing -> (0,0,0)), B-factor of CA, trailing OXT
// The host parser parsePdbChain (fcz_db.cpp) does the same on one core per entry; the two are tested against each other
// and against the reference's CPython module (tests/test_parse.py, tests/test_gpu_parse.py).
//
// Numeric fields: std::stof(field) = strtof.  A field of the shape every PDB writer produces -- blanks, optional sign,
// at most nine digits with at most one '.' -- is converted EXACTLY as strtof would (one correctly rounded double
// division; a quotient that lands on a float midpoint is resolved in integer arithmetic).  Any other field (exponents,
// hex floats, "nan", ten or more digits) makes the entry FCZ_E_ARG: strtof's full grammar is not reproduced.
#ifndef FCZ_PARSE_H
#define FCZ_PARSE_H

#include "fcz_format.h"

namespace fcz {

struct ParseTables {
    uint32_t atom[FCZ_NUM_CODES][FCZ_MAX_ATOMS];  // atom names of the table slots, up to four characters packed little-endian
    uint32_t res3[FCZ_NUM_CODES];
    uint8_t natoms[FCZ_NUM_CODES];
    uint32_t ca, oxt, n;
};
inline uint32_t name_key4(const char* z) {
    uint32_t k = 0;
    for (int i = 0; i < 4 && z[i]; i++) k |= (uint32_t)(uint8_t)z[i] << (8 * i);
    return k;
}
inline void build_parse_tables(ParseTables* t) {
    for (int c = 0; c < FCZ_NUM_CODES; c++) {
        t->res3[c] = name_key4(FCZ_NAME3[c]);
        t->natoms[c] = FCZ_NATOMS[c];
        for (int k = 0; k < FCZ_MAX_ATOMS; k++) t->atom[c][k] = name_key4(FCZ_ATOM_NAME[c][k]);
    }
    t->ca = name_key4("CA");
    t->oxt = name_key4("OXT");
    t->n = name_key4("N");
}

// trim(" \t") of the reference, then the first four characters as one integer key
FCZ_HD uint32_t parse_trim_key(const char* s, uint32_t n) {
    uint32_t a = 0, b = n;
    while (a < b && (s[a] == ' ' || s[a] == '\t')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) b--;
    uint32_t k = 0;
    for (uint32_t i = a; i < b && i < a + 4u; i++) k |= (uint32_t)(uint8_t)s[i] << (8u * (i - a));
    return k;
}
FCZ_HD int parse_int_field(const char* s, uint32_t n) {  // std::stoi on the field: blanks, sign, digits
    uint32_t i = 0;
    while (i < n && (s[i] == ' ' || s[i] == '\t')) i++;
    bool neg = false;
    if (i < n && (s[i] == '-' || s[i] == '+')) { neg = s[i] == '-'; i++; }
    long long v = 0;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) v = v * 10 + (s[i] - '0');
    return (int)(neg ? -v : v);
}
// strtof of a plain fixed-point field; *ok = false when the field has another shape (see the header comment)
FCZ_HD float parse_fixed_float(const char* s, uint32_t n, bool* ok) {
    uint32_t j = 0;
    while (j < n && (s[j] == ' ' || s[j] == '\t')) j++;
    bool neg = false;
    if (j < n && (s[j] == '-' || s[j] == '+')) { neg = s[j] == '-'; j++; }
    uint64_t m = 0;
    int nd = 0, frac = 0;
    bool dot = false, good = true;
    for (; j < n; j++) {
        const char ch = s[j];
        if (ch >= '0' && ch <= '9') { m = m * 10u + (uint64_t)(ch - '0'); nd++; if (dot) frac++; }
        else if (ch == '.' && !dot) dot = true;
        else if (ch == ' ' || ch == '\t' || ch == '\r' || ch == 0) break;  // strtof stops here
        else { good = false; break; }
    }
    if (!good || nd == 0 || nd > 9) { *ok = false; return 0.0f; }
    double p = 1.0;
    for (int i = 0; i < frac; i++) p *= 10.0;  // exact: frac <= 9
    const double d = (double)m / p;            // one correctly rounded division of two exact doubles
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(d);
#else
    memcpy(&bits, &d, 8);
#endif
    float f = (float)d;
    if ((bits & 0x1FFFFFFFull) == 0x10000000ull && m != 0) {
        // d sits exactly on the midpoint of two floats, possibly only because the division was rounded: decide from the
        // exact sign of m / 10^frac - d.  d = mant * 2^e2 with a 25-bit integer mant; compare m * 2^(-e2) with
        // mant * 10^frac in integers.  Above the midpoint -> the upper float, below -> the lower one; an exact tie (the
        // decimal IS the midpoint) rounds to even, which (float)d already did.
        const int e2 = (int)((bits >> 52) & 0x7FF) - 1075 + 28;
        const uint64_t mant = ((bits & 0xFFFFFFFFFFFFFull) | (1ull << 52)) >> 28;
        uint64_t p10 = 1;
        for (int i = 0; i < frac; i++) p10 *= 10u;
        unsigned __int128 lhs = (unsigned __int128)m, rhs = (unsigned __int128)mant * p10;
        if (e2 >= 0 && e2 < 40) rhs <<= e2; else if (e2 < 0 && e2 > -90) lhs <<= -e2;
        if (lhs != rhs) {
            const uint32_t fb = f2u(f);
            const bool f_above = (double)f > d;
            const float below = f_above ? u2f(fb - 1u) : f, above = f_above ? f : u2f(fb + 1u);
            f = lhs > rhs ? above : below;
        }
    }
    *ok = true;
    return neg ? -f : f;
}

struct RawAtom {
    uint32_t name;  // trimmed atom name, packed
    uint32_t res;   // trimmed residue name, packed
    int32_t serial, resnum;
    float x, y, z, b;
};

// One entry's text and its workspace.  `lines` / `raw` hold up to max_lines entries.
struct ParseEntry {
    const char* text;
    uint32_t len;
    uint32_t max_lines;
    uint32_t* lines;     // [max_lines + 1] start offset of every line (after the count pass: of every ATOM line kept)
    RawAtom* raw;        // [max_lines] parsed ATOM lines, alternative positions already dropped
    uint32_t* rstart;    // [max_lines + 1] first kept atom of every residue
    uint32_t* scratch;   // [8] words visible to all threads of the context: flags, counts
};
enum { PS_FLAG = 0, PS_NATOM = 1, PS_NRES = 2, PS_NSLOT = 3, PS_CHAIN = 4, PS_NLINES = 5, PS_BITS = 6 };

FCZ_HD int parse_code_of(const ParseTables* pt, uint32_t key) {
    for (int c = 0; c < FCZ_NUM_CODES; c++)
        if (pt->res3[c] == key) return pt->natoms[c] ? c : FCZ_CODE_UNK;
    return FCZ_CODE_UNK;
}

// Plan of one entry: finds the ATOM lines, parses them, drops alternative positions, splits residues.  Leaves the kept
// atoms in e.raw, the residue starts in e.rstart and in e.scratch: [PS_FLAG] 0 or the parser's flag (1 no ATOM line,
// 2 several chains, 3 short ATOM line, 4 a numeric field of another shape), [PS_NATOM] kept atoms, [PS_NRES] residues,
// [PS_NSLOT] table slots (= atoms of the canonical layout).  Every thread of the context must call it.
template <class Ctx>
FCZ_HD void parse_entry_plan(Ctx& cx, const ParseTables* pt, const ParseEntry& e) {
    const uint32_t len = e.len;
    if (cx.tid == 0) { for (int i = 0; i < 8; i++) e.scratch[i] = 0u; e.scratch[PS_CHAIN] = 0xFFFFFFFFu; }
    cx.sync();
    // ---- pass 1: line starts.  Every thread owns a contiguous byte range; a line belongs to the thread that holds its
    // first byte.
    const uint32_t chunk = (len + (uint32_t)cx.nthr - 1u) / (uint32_t)cx.nthr;
    uint32_t b0 = (uint32_t)cx.tid * chunk; if (b0 > len) b0 = len;
    uint32_t b1 = b0 + chunk; if (b1 > len) b1 = len;
    uint32_t cnt = 0;
    for (uint32_t i = b0; i < b1; i++) cnt += (i == 0u || e.text[i - 1u] == '\n') ? 1u : 0u;
    uint32_t base = cx.excl_scan(cnt);
    const uint32_t my_first = base;
    for (uint32_t i = b0; i < b1; i++)
        if (i == 0u || e.text[i - 1u] == '\n') { if (base < e.max_lines) e.lines[base] = i; base++; }
    if (b1 == len && cx.tid == cx.nthr - 1) e.scratch[PS_NLINES] = base;  // number of lines (the last thread's range ends the text)
    (void)my_first;
    cx.sync();
    const uint32_t n_lines = e.scratch[PS_NLINES];
    if (n_lines > e.max_lines) { if (cx.tid == 0) e.scratch[PS_FLAG] = 3u; cx.sync(); return; }
    // ---- pass 2: ATOM lines -> RawAtom at the line's index (holes for other records), flags
    uint32_t flag = 0;
    for (uint32_t l = (uint32_t)cx.tid; l < n_lines; l += (uint32_t)cx.nthr) {
        const uint32_t s = e.lines[l];
        uint32_t n = (l + 1u < n_lines ? e.lines[l + 1u] - 1u : len) - s;  // without the '\n'
        if (l + 1u == n_lines && n > 0u && e.text[s + n - 1u] == '\n') n--;
        const char* line = e.text + s;
        RawAtom a;
        a.name = 0xFFFFFFFFu;  // not an ATOM line
        if (n >= 4u && line[0] == 'A' && line[1] == 'T' && line[2] == 'O' && line[3] == 'M') {
            if (n < 61u) { flag |= 4u; }  // substr(21,1) / the B-factor column would throw in the reference (flag 3)
            else {
                bool o1, o2, o3, o4;
                a.name = parse_trim_key(line + 12, 4u);
                a.res = parse_trim_key(line + 17, 3u);
                a.serial = parse_int_field(line + 6, 5u);
                a.resnum = parse_int_field(line + 22, 4u);
                a.x = parse_fixed_float(line + 30, 8u, &o1);
                a.y = parse_fixed_float(line + 38, 8u, &o2);
                a.z = parse_fixed_float(line + 46, 8u, &o3);
                a.b = parse_fixed_float(line + 60, n - 60u < 6u ? n - 60u : 6u, &o4);
                if (!(o1 && o2 && o3 && o4)) flag |= 8u;
                // chain of the FIRST ATOM line: the lowest line index wins (atomic min on line << 8 | chain)
                cx.atomic_min_u(&e.scratch[PS_CHAIN], (l << 8) | (uint32_t)(uint8_t)line[21]);
                // the chain character rides in the top byte of `res` (names are at most three characters)
                a.res = (a.res & 0x00FFFFFFu) | ((uint32_t)(uint8_t)line[21] << 24);
            }
        }
        e.raw[l] = a;
    }
    if (flag) cx.atomic_or_u(&e.scratch[PS_BITS], flag);
    cx.sync();
    const uint32_t first_chain = e.scratch[PS_CHAIN] & 0xFFu;
    const bool none = e.scratch[PS_CHAIN] == 0xFFFFFFFFu;
    // ---- pass 3: keep = ATOM line whose name differs from the PREVIOUS ATOM line's (removeAlternativePosition compares
    // with the last kept atom, which carries the same name as the previous line whether that one was kept or not);
    // compaction of the kept atoms in place, in order
    const uint32_t lchunk = (n_lines + (uint32_t)cx.nthr - 1u) / (uint32_t)cx.nthr;
    uint32_t l0 = (uint32_t)cx.tid * lchunk; if (l0 > n_lines) l0 = n_lines;
    uint32_t l1 = l0 + lchunk; if (l1 > n_lines) l1 = n_lines;
    // name of the last ATOM line before l0
    uint32_t prev = 0xFFFFFFFEu;
    for (uint32_t l = l0; l > 0u; l--) { const uint32_t nm = e.raw[l - 1u].name; if (nm != 0xFFFFFFFFu) { prev = nm; break; } }
    uint32_t kept = 0, multi = 0;
    {
        uint32_t pv = prev;
        for (uint32_t l = l0; l < l1; l++) {
            const uint32_t nm = e.raw[l].name;
            if (nm == 0xFFFFFFFFu) continue;
            if ((e.raw[l].res >> 24) != first_chain) multi = 1u;
            if (nm != pv) kept++;
            pv = nm;
        }
    }
    if (multi) cx.atomic_or_u(&e.scratch[PS_BITS], 2u);
    cx.sync();  // every thread has read its predecessors' names before anything moves
    uint32_t kbase = cx.excl_scan(kept);
    // the compaction writes to indices <= the source index, but another thread's sources may sit there: stage through
    // registers is not possible for arbitrary counts, so kept atoms go to the `lines` array as indices first
    {
        uint32_t pv = prev, k = kbase;
        for (uint32_t l = l0; l < l1; l++) {
            const uint32_t nm = e.raw[l].name;
            if (nm == 0xFFFFFFFFu) continue;
            if (nm != pv) e.rstart[k++] = l;  // rstart doubles as the kept-line index list for a moment
            pv = nm;
        }
        if (l1 == n_lines && cx.tid == cx.nthr - 1) e.scratch[PS_NATOM] = k;
    }
    cx.sync();
    const uint32_t n_atoms = e.scratch[PS_NATOM];
    const uint32_t bits = e.scratch[PS_BITS];
    if (none || n_atoms == 0u || bits) {
        if (cx.tid == 0) e.scratch[PS_FLAG] = (bits & 4u) ? 3u : ((bits & 2u) ? 2u : ((bits & 8u) ? 4u : 1u));
        cx.sync();
        return;
    }
    // gather the kept atoms into a dense list: lines[] is free now (line starts are no longer needed) -- it becomes
    // the list of kept line indices, and raw[] is compacted through it in ascending order by ONE pass per thread
    // range after a barrier (a kept atom only moves DOWN, to an index no thread reads later than it writes: see below)
    for (uint32_t k = (uint32_t)cx.tid; k < n_atoms; k += (uint32_t)cx.nthr) e.lines[k] = e.rstart[k];
    cx.sync();
    // moving raw[lines[k]] -> raw[k] in parallel is a hazard (k may be another thread's source), so the dense copy
    // lives in the upper half of the workspace: raw2 = raw + max_lines (the caller provides 2 * max_lines RawAtoms)
    RawAtom* dense = e.raw + e.max_lines;
    for (uint32_t k = (uint32_t)cx.tid; k < n_atoms; k += (uint32_t)cx.nthr) dense[k] = e.raw[e.lines[k]];
    cx.sync();
    // ---- pass 4: residue starts (splitAtomByResidue): atom j opens a residue when j == 0, or its residue number
    // differs from its predecessor's and it is not the last atom
    const uint32_t achunk = (n_atoms + (uint32_t)cx.nthr - 1u) / (uint32_t)cx.nthr;
    uint32_t a0 = (uint32_t)cx.tid * achunk; if (a0 > n_atoms) a0 = n_atoms;
    uint32_t a1 = a0 + achunk; if (a1 > n_atoms) a1 = n_atoms;
    uint32_t nst = 0;
    for (uint32_t j = a0; j < a1; j++) nst += (j == 0u || (dense[j].resnum != dense[j - 1u].resnum && j != n_atoms - 1u)) ? 1u : 0u;
    uint32_t rbase = cx.excl_scan(nst);
    for (uint32_t j = a0; j < a1; j++)
        if (j == 0u || (dense[j].resnum != dense[j - 1u].resnum && j != n_atoms - 1u)) e.rstart[rbase++] = j;
    if (a1 == n_atoms && cx.tid == cx.nthr - 1) { e.scratch[PS_NRES] = rbase; e.rstart[rbase] = n_atoms; }
    // fragments (identifyDiscontinousResInd, src/atom_coordinate.cpp:506-530; src/main.cpp:469-484): the reference CLI cuts a
    // chain where the residue number of an N atom exceeds that of the previous N atom by more than one, and starts its
    // first fragment at the chain's first N atom.  Such an entry is not ONE chain for `foldcomp compress`: flag 5 (the batch
    // front end splits it on the host, fczgpu::parsePdbUnits).
    {
        uint32_t gap = 0;
        bool have = false;
        int32_t prevn = 0;
        for (uint32_t j = a0; j > 0u && a0 < a1; j--)  // the last N atom before this thread's range
            if (dense[j - 1u].name == pt->n) { have = true; prevn = dense[j - 1u].resnum; break; }
        for (uint32_t j = a0; j < a1; j++) {
            if (dense[j].name != pt->n) continue;
            if (have && dense[j].resnum - prevn > 1) gap = 1u;
            have = true; prevn = dense[j].resnum;
        }
        if (cx.tid == 0 && dense[0].name != pt->n) gap = 1u;
        if (gap) cx.atomic_or_u(&e.scratch[PS_BITS], 16u);
    }
    cx.sync();
    if (e.scratch[PS_BITS] & 16u) {
        if (cx.tid == 0) e.scratch[PS_FLAG] = 5u;
        cx.sync();
        return;
    }
    // ---- pass 5: table slots of every residue (the canonical atom count)
    const uint32_t n_res = e.scratch[PS_NRES];
    uint32_t slots = 0;
    for (uint32_t r = (uint32_t)cx.tid; r < n_res; r += (uint32_t)cx.nthr) slots += pt->natoms[parse_code_of(pt, dense[e.rstart[r]].res & 0x00FFFFFFu)];
    if (slots) cx.atomic_add(&e.scratch[PS_NSLOT], slots);
    cx.sync();
}

// Emit of one planned entry into the canonical layout (the arrays of ONE chain): residue codes, B-factors, slots, meta.
template <class Ctx>
FCZ_HD void parse_entry_emit(Ctx& cx, const ParseTables* pt, const ParseEntry& e, uint8_t* res_type, float* bfactor, float* xyz,
                             fcz_chain_meta* meta) {
    const RawAtom* dense = e.raw + e.max_lines;
    const uint32_t n_atoms = e.scratch[PS_NATOM], n_res = e.scratch[PS_NRES];
    // atom offset of every residue: exclusive scan of the table atom counts
    const uint32_t chunk = (n_res + (uint32_t)cx.nthr - 1u) / (uint32_t)cx.nthr;
    uint32_t r0 = (uint32_t)cx.tid * chunk; if (r0 > n_res) r0 = n_res;
    uint32_t r1 = r0 + chunk; if (r1 > n_res) r1 = n_res;
    uint32_t sum = 0;
    for (uint32_t r = r0; r < r1; r++) sum += pt->natoms[parse_code_of(pt, dense[e.rstart[r]].res & 0x00FFFFFFu)];
    uint32_t base = cx.excl_scan(sum);
    for (uint32_t r = r0; r < r1; r++) {
        const uint32_t i = e.rstart[r], j = e.rstart[r + 1u];
        const int code = parse_code_of(pt, dense[i].res & 0x00FFFFFFu);
        res_type[r] = (uint8_t)code;
        const uint32_t na = pt->natoms[code];
        for (uint32_t k = 0; k < na; k++) {
            const uint32_t want = pt->atom[code][k];
            float x = 0.f, y = 0.f, z = 0.f;  // findFirstAtomCoords: a missing atom reads as (0,0,0)
            for (uint32_t a = i; a < j; a++)
                if (dense[a].name == want) { x = dense[a].x; y = dense[a].y; z = dense[a].z; break; }
            float* o = xyz + 3u * (base + k);
            o[0] = x; o[1] = y; o[2] = z;
        }
        float bf = 0.f;
        for (uint32_t a = i; a < j; a++)
            if (dense[a].name == pt->ca) { bf = dense[a].b; break; }
        bfactor[r] = bf;
        base += na;
    }
    if (cx.tid == 0) {
        fcz_chain_meta m;
        m.n_atom = (uint16_t)n_atoms;
        m.idx_residue = (uint16_t)dense[0].resnum;
        m.idx_atom = (uint16_t)dense[0].serial;
        m.chain = (uint8_t)(dense[0].res >> 24);
        m.has_oxt = 0; m.oxt[0] = 0.f; m.oxt[1] = 0.f; m.oxt[2] = 0.f;
        if (dense[n_atoms - 1u].name == pt->oxt) {  // src/foldcomp.cpp:473-481
            m.has_oxt = 1;
            m.oxt[0] = dense[n_atoms - 1u].x; m.oxt[1] = dense[n_atoms - 1u].y; m.oxt[2] = dense[n_atoms - 1u].z;
        }
        *meta = m;
    }
}

}  // namespace fcz
#endif  // FCZ_PARSE_H
