// foldcomp_b200/csrc/fcz_text.h -- PDB text of decoded chains and the FCZ "extract" scans, written once for host
// and device (the CUDA kernels in fcz_engine.cu and the one-thread CPU model in tests/emu/ compile this file).
//
// Reference code replaced:
//   writeAtomCoordinatesToPDB   src/atom_coordinate.cpp:220-291   TITLE lines, one ATOM line per atom, TER
//   fast_ftoa<T,P>              src/atom_coordinate.cpp:186-218   the reference's own float formatter
//   Foldcomp::extract           src/foldcomp.cpp:1260-1336        pLDDT digits / amino-acid sequence of a blob
//
// An ATOM line is 81 bytes when every field fits its column ("uniform"); std::setw is a MINIMUM width, so a
// serial > 99999, a residue number > 9999, a coordinate outside (-999.9995, 9999.9995) or a B-factor outside
// (-99.995, 999.995) makes the line longer.  Both cases are reproduced byte for byte; the uniform case is the
// fast path (line offset = 81 * atom index).
#ifndef FCZ_TEXT_H
#define FCZ_TEXT_H

#include <stdint.h>

#include "../../include/fcz_engine.h"
#include "fcz_format.h"
#include "fcz_math.h"
#include "fcz_tables.h"

namespace fcz {

#define FCZ_PDB_LINE 81u  // bytes of a uniform ATOM line
#define FCZ_PDB_TER 27u   // bytes of a uniform TER line

// Names as the emitter needs them: residue names and atom names packed little-endian into 32-bit words, atom
// names already left-justified in 3 columns (std::setw(3) << std::left, src/atom_coordinate.cpp:253-255).
struct TextTables {  // FCZ_CODE_ROWS rows: 24..31 repeat UNK (see Tables in fcz_codec.h)
    uint32_t name3[FCZ_CODE_ROWS];                 // 'A' | 'L' << 8 | 'A' << 16
    uint32_t atom[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];   // "CA " etc; byte 0 is also the element column (atom[0])
    uint8_t natoms[FCZ_CODE_ROWS];
    uint8_t alt[FCZ_CODE_ROWS][FCZ_MAX_ATOMS];
    uint8_t name1[FCZ_CODE_ROWS];
};

inline void build_text_tables(TextTables* t) {
    for (int row = 0; row < FCZ_CODE_ROWS; row++) {
        const int c = (int)norm_code((unsigned)row);
        t->name3[row] = (uint32_t)(uint8_t)FCZ_NAME3[c][0] | (uint32_t)(uint8_t)FCZ_NAME3[c][1] << 8 | (uint32_t)(uint8_t)FCZ_NAME3[c][2] << 16;
        t->natoms[row] = FCZ_NATOMS[c];
        t->name1[row] = (uint8_t)FCZ_NAME1[c];
        for (int k = 0; k < FCZ_MAX_ATOMS; k++) {
            uint32_t w = 0;
            bool end = false;
            for (int j = 0; j < 3; j++) {
                char ch = end ? ' ' : FCZ_ATOM_NAME[c][k][j];
                if (ch == 0) { end = true; ch = ' '; }
                w |= (uint32_t)(uint8_t)ch << (8 * j);
            }
            t->atom[row][k] = w;
            t->alt[row][k] = FCZ_ALT[c][k];
        }
    }
}

// ---- fast_ftoa<T,P> (src/atom_coordinate.cpp:186-218), split into its two integers.  rounded = n -+ 0.5f/T,
// integer = (int32)rounded, decimal = (int32)((rounded - (float)integer) * T), both in FLOAT arithmetic without
// contraction; a negative n prints '-' followed by the absolute values (so -0.0004 prints "-0.000").
struct FtoaParts {
    uint32_t ip, dp;  // integer part, decimal part (absolute values)
    uint32_t neg;     // 1: leading '-'
};
FCZ_HD FtoaParts ftoa_parts(float n, float T, float half) {
    FtoaParts p;
    p.neg = n < 0 ? 1u : 0u;
    const float rounded = n + (p.neg ? -half : half);
    // |rounded| >= 2^31 (or NaN) is undefined behaviour in the reference; here it saturates
    const float rc = rounded >= 2147483520.0f ? 2147483520.0f : (rounded <= -2147483520.0f ? -2147483520.0f : (rounded == rounded ? rounded : 0.0f));
    const int32_t integer = (int32_t)rc;
    const float fr = (rc - (float)integer) * T;
    const int32_t decimal = (int32_t)fr;
    p.ip = (uint32_t)(integer < 0 ? -integer : integer);
    p.dp = (uint32_t)(decimal < 0 ? -decimal : decimal);
    if (!p.neg) { p.ip = (uint32_t)integer; p.dp = (uint32_t)decimal; }  // the reference takes abs() only when n < 0
    return p;
}
FCZ_HD uint32_t dec_digits(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10u) { v /= 10u; d++; }
    return d;
}
// characters fast_ftoa<T,P> produces
FCZ_HD uint32_t ftoa_len(const FtoaParts& p, uint32_t P) {
    const uint32_t dd = dec_digits(p.dp);
    return p.neg + dec_digits(p.ip) + 1u + (dd > P ? dd : P);
}
// write v as exactly nd decimal digits ending just before `end`; returns the first written position
FCZ_HD char* put_digits_before(char* end, uint32_t v, uint32_t nd) {
    for (uint32_t i = 0; i < nd; i++) { *--end = (char)('0' + v % 10u); v /= 10u; }
    return end;
}
// fast_ftoa<T,P> right-aligned in a field of `width` columns starting at dst (longer when it does not fit);
// returns the number of bytes written
FCZ_HD uint32_t put_ftoa(char* dst, const FtoaParts& p, uint32_t P, uint32_t width) {
    const uint32_t len = ftoa_len(p, P);
    const uint32_t w = len > width ? len : width;
    char* e = dst + w;
    const uint32_t dd = dec_digits(p.dp);
    e = put_digits_before(e, p.dp, dd > P ? dd : P);
    *--e = '.';
    e = put_digits_before(e, p.ip, dec_digits(p.ip));
    if (p.neg) *--e = '-';
    while (e > dst) *--e = ' ';
    return w;
}
// unsigned integer right-aligned in `width` columns (longer when it does not fit)
FCZ_HD uint32_t put_uint(char* dst, uint32_t v, uint32_t width) {
    const uint32_t nd = dec_digits(v);
    const uint32_t w = nd > width ? nd : width;
    char* e = put_digits_before(dst + w, v, nd);
    while (e > dst) *--e = ' ';
    return w;
}

struct AtomRec {       // everything one ATOM line shows
    uint32_t serial;   // AtomCoordinate::atom_index
    uint32_t resnum;   // AtomCoordinate::residue_index
    uint32_t name;     // packed atom name, 3 columns left-justified
    uint32_t res3;     // packed residue name
    uint8_t chain;
    float x, y, z, b;
};

// bytes beyond 81 that this atom's line needs
FCZ_HD uint32_t atom_line_extra(const AtomRec& a) {
    uint32_t ex = 0;
    uint32_t d = dec_digits(a.serial);
    ex += d > 5u ? d - 5u : 0u;
    d = dec_digits(a.resnum);
    ex += d > 4u ? d - 4u : 0u;
    const float h3 = 0.5f / 1000.0f, h2 = 0.5f / 100.0f;
    uint32_t l = ftoa_len(ftoa_parts(a.x, 1000.0f, h3), 3u);
    ex += l > 8u ? l - 8u : 0u;
    l = ftoa_len(ftoa_parts(a.y, 1000.0f, h3), 3u);
    ex += l > 8u ? l - 8u : 0u;
    l = ftoa_len(ftoa_parts(a.z, 1000.0f, h3), 3u);
    ex += l > 8u ? l - 8u : 0u;
    l = ftoa_len(ftoa_parts(a.b, 100.0f, h2), 2u);
    ex += l > 6u ? l - 6u : 0u;
    return ex;
}

// ---- the uniform (81-byte) line as 21 little-endian words: no loops, no per-character stores.
// A line is uniform iff every field fits its column; that is decided from the same integers the formatter prints.
FCZ_HD bool ftoa_fits(const FtoaParts& p, uint32_t ip_max_pos, uint32_t ip_max_neg, uint32_t dp_max) {
    return p.ip <= (p.neg ? ip_max_neg : ip_max_pos) && p.dp <= dp_max;
}
FCZ_HD bool atom_line_uniform(const AtomRec& a) {
    const float h3 = 0.5f / 1000.0f, h2 = 0.5f / 100.0f;
    return a.serial <= 99999u && a.resnum <= 9999u && ftoa_fits(ftoa_parts(a.x, 1000.0f, h3), 9999u, 999u, 999u) &&
           ftoa_fits(ftoa_parts(a.y, 1000.0f, h3), 9999u, 999u, 999u) && ftoa_fits(ftoa_parts(a.z, 1000.0f, h3), 9999u, 999u, 999u) &&
           ftoa_fits(ftoa_parts(a.b, 100.0f, h2), 999u, 99u, 99u);
}
// v <= 9999 as four characters, right-aligned, blank-padded, with a '-' before the first digit when neg (v <= 999 then)
FCZ_HD uint32_t digits4_right(uint32_t v, uint32_t neg) {
    const uint32_t d3 = v / 1000u, r3 = v - d3 * 1000u, d2 = r3 / 100u, r2 = r3 - d2 * 100u, d1 = r2 / 10u, d0 = r2 - d1 * 10u;
    const uint32_t D = 0x30303030u + (d3 | d2 << 8 | d1 << 16 | d0 << 24);
    const uint32_t nd = 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u);
    const uint32_t shift = 8u * (4u - nd);
    const uint32_t mask = 0xFFFFFFFFu << shift;
    uint32_t out = (D & mask) | (0x20202020u & ~mask);
    if (neg) out ^= (uint32_t)(' ' ^ '-') << (shift - 8u);
    return out;
}
// columns (0-based): 0-5 "ATOM  ", 6-10 serial, 11-12 blanks, 13-15 name, 16 blank, 17-19 residue, 20 blank, 21 chain,
// 22-25 residue number, 26-29 blanks, 30-37 x, 38-45 y, 46-53 z, 54-59 "  1.00", 60-65 B, 66-76 blanks, 77 element,
// 78-79 blanks, 80 newline; word j holds columns 4j..4j+3
FCZ_HD void atom_line_words(uint32_t* w, const AtomRec& a) {
    const float h3 = 0.5f / 1000.0f, h2 = 0.5f / 100.0f;
    w[0] = (uint32_t)'A' | (uint32_t)'T' << 8 | (uint32_t)'O' << 16 | (uint32_t)'M' << 24;
    {
        const uint32_t hi = a.serial / 10000u, lo = a.serial - hi * 10000u;
        uint32_t c0 = ' ', R;
        if (hi) {
            const uint32_t d3 = lo / 1000u, r3 = lo - d3 * 1000u, d2 = r3 / 100u, r2 = r3 - d2 * 100u, d1 = r2 / 10u, d0 = r2 - d1 * 10u;
            c0 = '0' + hi;
            R = 0x30303030u + (d3 | d2 << 8 | d1 << 16 | d0 << 24);
        } else {
            R = digits4_right(lo, 0u);
        }
        w[1] = 0x00002020u | c0 << 16 | (R & 0xFFu) << 24;
        w[2] = (R >> 8) | (uint32_t)' ' << 24;
    }
    w[3] = (uint32_t)' ' | (a.name & 0xFFFFFFu) << 8;
    w[4] = (uint32_t)' ' | (a.res3 & 0xFFFFFFu) << 8;
    {
        const uint32_t rn = digits4_right(a.resnum, 0u);
        w[5] = (uint32_t)' ' | (uint32_t)a.chain << 8 | (rn & 0xFFFFu) << 16;
        w[6] = (rn >> 16) | 0x20200000u;
    }
    uint32_t fl[3], fh[3];
    const float xyz[3] = {a.x, a.y, a.z};
    for (int k = 0; k < 3; k++) {
        const FtoaParts p = ftoa_parts(xyz[k], 1000.0f, h3);
        fl[k] = digits4_right(p.ip, p.neg);
        const uint32_t e2 = p.dp / 100u, r2 = p.dp - e2 * 100u, e1 = r2 / 10u, e0 = r2 - e1 * 10u;
        fh[k] = (uint32_t)'.' | ('0' + e2) << 8 | ('0' + e1) << 16 | ('0' + e0) << 24;
    }
    w[7] = 0x00002020u | fl[0] << 16;
    w[8] = fl[0] >> 16 | fh[0] << 16;
    w[9] = fh[0] >> 16 | fl[1] << 16;
    w[10] = fl[1] >> 16 | fh[1] << 16;
    w[11] = fh[1] >> 16 | fl[2] << 16;
    w[12] = fl[2] >> 16 | fh[2] << 16;
    w[13] = fh[2] >> 16 | 0x20200000u;
    w[14] = (uint32_t)'1' | (uint32_t)'.' << 8 | (uint32_t)'0' << 16 | (uint32_t)'0' << 24;
    {
        const FtoaParts p = ftoa_parts(a.b, 100.0f, h2);
        const uint32_t bi = digits4_right(p.ip, p.neg) >> 8;  // three columns
        const uint32_t e1 = p.dp / 10u, e0 = p.dp - e1 * 10u;
        w[15] = bi | (uint32_t)'.' << 24;
        w[16] = ('0' + e1) | ('0' + e0) << 8 | 0x20200000u;
    }
    w[17] = 0x20202020u;
    w[18] = 0x20202020u;
    w[19] = (uint32_t)' ' | (a.name & 0xFFu) << 8 | 0x20200000u;
    w[20] = (uint32_t)'\n';
}
// the low 32 bits of (hi:lo) >> s, 0 < s < 32
FCZ_HD uint32_t shr64(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
#endif
}
// store the 81 bytes held in w[0..20] at dst (any alignment): aligned 32-bit stores for the interior, bytes at the ends
FCZ_HD void store_line81(char* dst, const uint32_t* w) {
    const uint32_t m = (uint32_t)((uintptr_t)dst & 3u);
    if (m == 0u) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int k = 0; k < 20; k++) d[k] = w[k];
        dst[80] = '\n';
        return;
    }
    const uint32_t s = 8u * (4u - m);  // aligned word k (k >= 1) starts at line byte 4k - m
    uint32_t* d = reinterpret_cast<uint32_t*>(dst - m);
    for (uint32_t b = 0; b < 4u - m; b++) dst[b] = (char)(w[0] >> (8u * b));
    for (int k = 1; k < 20; k++) d[k] = shr64(w[k - 1], w[k], s);
    const uint32_t tail = shr64(w[19], w[20], s);  // line bytes 80-m .. 80 (m + 1 of them)
    for (uint32_t b = 0; b <= m; b++) dst[80u - m + b] = (char)(tail >> (8u * b));
}

// One ATOM line (src/atom_coordinate.cpp:246-275); returns its length.
FCZ_HD uint32_t put_atom_line(char* dst, const AtomRec& a) {
    char* p = dst;
    p[0] = 'A'; p[1] = 'T'; p[2] = 'O'; p[3] = 'M'; p[4] = ' '; p[5] = ' ';
    p += 6;
    p += put_uint(p, a.serial, 5u);
    *p++ = ' ';
    *p++ = ' ';  // names in the tables have at most 3 characters: " " + setw(3) left
    *p++ = (char)(a.name & 0xFFu); *p++ = (char)((a.name >> 8) & 0xFFu); *p++ = (char)((a.name >> 16) & 0xFFu);
    *p++ = ' ';
    *p++ = (char)(a.res3 & 0xFFu); *p++ = (char)((a.res3 >> 8) & 0xFFu); *p++ = (char)((a.res3 >> 16) & 0xFFu);
    *p++ = ' ';
    *p++ = (char)a.chain;
    p += put_uint(p, a.resnum, 4u);
    p[0] = ' '; p[1] = ' '; p[2] = ' '; p[3] = ' ';
    p += 4;
    const float h3 = 0.5f / 1000.0f, h2 = 0.5f / 100.0f;
    p += put_ftoa(p, ftoa_parts(a.x, 1000.0f, h3), 3u, 8u);
    p += put_ftoa(p, ftoa_parts(a.y, 1000.0f, h3), 3u, 8u);
    p += put_ftoa(p, ftoa_parts(a.z, 1000.0f, h3), 3u, 8u);
    p[0] = ' '; p[1] = ' '; p[2] = '1'; p[3] = '.'; p[4] = '0'; p[5] = '0';
    p += 6;
    p += put_ftoa(p, ftoa_parts(a.b, 100.0f, h2), 2u, 6u);
    for (int i = 0; i < 10; i++) *p++ = ' ';
    *p++ = ' ';
    *p++ = (char)(a.name & 0xFFu);  // std::setw(2) << atom[0]
    *p++ = ' '; *p++ = ' '; *p++ = '\n';
    return (uint32_t)(p - dst);
}

// TER line after the last atom (src/atom_coordinate.cpp:276-287)
FCZ_HD uint32_t ter_line_len(const AtomRec& last) {
    const uint32_t d1 = dec_digits(last.serial + 1u), d2 = dec_digits(last.resnum);
    return FCZ_PDB_TER + (d1 > 5u ? d1 - 5u : 0u) + (d2 > 4u ? d2 - 4u : 0u);
}
FCZ_HD uint32_t put_ter_line(char* dst, const AtomRec& last) {
    char* p = dst;
    p[0] = 'T'; p[1] = 'E'; p[2] = 'R'; p[3] = ' '; p[4] = ' '; p[5] = ' ';
    p += 6;
    p += put_uint(p, last.serial + 1u, 5u);
    for (int i = 0; i < 6; i++) *p++ = ' ';
    *p++ = (char)(last.res3 & 0xFFu); *p++ = (char)((last.res3 >> 8) & 0xFFu); *p++ = (char)((last.res3 >> 16) & 0xFFu);
    *p++ = ' ';
    *p++ = (char)last.chain;
    p += put_uint(p, last.resnum, 4u);
    *p++ = '\n';
    return (uint32_t)(p - dst);
}

// TITLE lines (src/atom_coordinate.cpp:223-243): "TITLE     " + 70 characters, then "TITLE  %3d" continuation
// lines numbered from 2 (the "% 3d" flag only matters for numbers of three digits or more: a space is prepended).
FCZ_HD uint32_t title_cont_prefix(uint32_t k) {  // bytes of "TITLE  % 3d"
    const uint32_t d = dec_digits(k);
    return 7u + (d + 1u > 3u ? d + 1u : 3u);
}
FCZ_HD uint32_t title_lines_len(uint32_t T) {
    if (T == 0u) return 0u;
    uint32_t n = 10u + (T < 70u ? T : 70u) + 1u;
    uint32_t k = 2u;
    for (uint32_t done = 70u; done < T; done += 70u, k++) n += title_cont_prefix(k) + (T - done < 70u ? T - done : 70u) + 1u;
    return n;
}
FCZ_HD uint32_t put_title_lines(char* dst, const char* title, uint32_t T) {
    if (T == 0u) return 0u;
    char* p = dst;
    const char* h = "TITLE     ";
    for (int i = 0; i < 10; i++) *p++ = h[i];
    uint32_t n = T < 70u ? T : 70u;
    for (uint32_t i = 0; i < n; i++) *p++ = title[i];
    *p++ = '\n';
    uint32_t k = 2u;
    for (uint32_t done = 70u; done < T; done += 70u, k++) {
        for (int i = 0; i < 7; i++) *p++ = h[i];
        const uint32_t d = dec_digits(k);
        p += put_uint(p, k, d + 1u > 3u ? d + 1u : 3u);
        n = T - done < 70u ? T - done : 70u;
        for (uint32_t i = 0; i < n; i++) *p++ = title[done + i];
        *p++ = '\n';
    }
    return (uint32_t)(p - dst);
}

// ------------------------------------------------------------------ chain-level plan and emit (any context)
// Written against the same abstract execution context as fcz_codec.h (tid, nthr, sync(), excl_scan()): the CUDA
// kernels instantiate them with a CTA, tests/emu/ with one host thread.

#define FCZ_PDB_UNIT_RES 32u                                      // residues per emit unit (one CTA)
#define FCZ_PDB_UNIT_ATOMS (FCZ_PDB_UNIT_RES * FCZ_MAX_ATOMS)     // 448 atoms at most
#define FCZ_PDB_STAGE_BYTES (FCZ_PDB_UNIT_ATOMS * FCZ_PDB_LINE + 32u)

struct PdbChain {
    uint32_t L, A, title_len;
    const uint8_t* type;   // [L] residue codes
    const float* bfac;     // [L]
    const float* X;        // [3A] atoms in the decoder's output order
    const char* title;
    const fcz_chain_meta* meta;
    int use_alt;           // atoms are in the -a order: names follow TextTables::alt
    uint32_t* aoff;        // [L+1] first atom of each residue, relative to the chain
    uint32_t* toff;        // [L+1] text offset of each residue's first line, relative to the chain's text;
                           //       toff[L] = where the OXT and TER lines start
};

FCZ_HD AtomRec pdb_atom_rec(const TextTables* tt, const PdbChain& ch, uint32_t r, uint32_t k, uint32_t atom) {
    const unsigned code = ch.type[r] & 31u;  // tables have FCZ_CODE_ROWS rows
    AtomRec a;
    a.serial = (uint32_t)ch.meta->idx_atom + atom;
    a.resnum = (uint32_t)ch.meta->idx_residue + r;
    a.name = tt->atom[code][ch.use_alt ? tt->alt[code][k] : k];
    a.res3 = tt->name3[code];
    a.chain = ch.meta->chain;
    a.x = ch.X[3u * atom]; a.y = ch.X[3u * atom + 1u]; a.z = ch.X[3u * atom + 2u];
    a.b = ch.bfac[r];
    return a;
}
// the OXT record (Foldcomp::read, src/foldcomp.cpp:958-961: residue_index = nResidue) -- also the atom TER refers to
FCZ_HD AtomRec pdb_oxt_rec(const TextTables* tt, const PdbChain& ch) {
    AtomRec a;
    a.serial = (uint32_t)ch.meta->idx_atom + ch.A;
    a.resnum = ch.L;
    a.name = (uint32_t)'O' | (uint32_t)'X' << 8 | (uint32_t)'T' << 16;
    a.res3 = tt->name3[ch.type[ch.L - 1u] & 31u];
    a.chain = ch.meta->chain;
    a.x = ch.meta->oxt[0]; a.y = ch.meta->oxt[1]; a.z = ch.meta->oxt[2];
    a.b = ch.bfac[ch.L - 1u];
    return a;
}
FCZ_HD AtomRec pdb_last_rec(const TextTables* tt, const PdbChain& ch) {
    if (ch.meta->has_oxt) return pdb_oxt_rec(tt, ch);
    const uint32_t r = ch.L - 1u;
    return pdb_atom_rec(tt, ch, r, ch.A - 1u - ch.aoff[r], ch.A - 1u);
}

// Plan: fills ch.aoff / ch.toff; returns the chain's text bytes (same value in every thread).
// `scratch` is one word visible to all threads of the context.
template <class Ctx>
FCZ_HD uint32_t pdb_plan_chain(Ctx& cx, const TextTables* tt, const PdbChain& ch, uint32_t* scratch) {
    const uint32_t L = ch.L;
    if (L == 0u) return 0u;
    const uint32_t chunk = (L + cx.nthr - 1) / cx.nthr;
    uint32_t r0 = cx.tid * chunk; if (r0 > L) r0 = L;
    uint32_t r1 = r0 + chunk; if (r1 > L) r1 = L;
    uint32_t sum = 0;
    for (uint32_t r = r0; r < r1; r++) sum += natoms_packed(ch.type[r] & 31u);
    uint32_t base = cx.excl_scan(sum);
    const uint32_t a_first = base;
    for (uint32_t r = r0; r < r1; r++) { ch.aoff[r] = base; base += natoms_packed(ch.type[r] & 31u); }
    if (r1 == L) ch.aoff[L] = base;
    cx.sync();
    // bytes beyond 81 (zero for ordinary coordinates): the cheap fits-its-columns test per atom, the exact
    // measurement only for the atoms that fail it
    uint32_t extra = 0, atom = a_first;
    for (uint32_t r = r0; r < r1; r++) {
        const uint32_t n = natoms_packed(ch.type[r] & 31u);
        for (uint32_t k = 0; k < n; k++, atom++) {
            const AtomRec a = pdb_atom_rec(tt, ch, r, k, atom);
            if (!atom_line_uniform(a)) extra += atom_line_extra(a);
        }
    }
    uint32_t ebase = cx.excl_scan(extra);
    const uint32_t head = title_lines_len(ch.title_len);
    atom = a_first;
    if (extra == 0u) {  // this thread's residues are uniform: offsets follow from the atom offsets alone
        for (uint32_t r = r0; r < r1; r++) { ch.toff[r] = head + FCZ_PDB_LINE * atom + ebase; atom += natoms_packed(ch.type[r] & 31u); }
    } else {
        for (uint32_t r = r0; r < r1; r++) {
            ch.toff[r] = head + FCZ_PDB_LINE * atom + ebase;
            const uint32_t n = natoms_packed(ch.type[r] & 31u);
            for (uint32_t k = 0; k < n; k++, atom++) ebase += atom_line_extra(pdb_atom_rec(tt, ch, r, k, atom));
        }
    }
    if (r1 == L) {
        const uint32_t t_tail = head + FCZ_PDB_LINE * atom + ebase;
        ch.toff[L] = t_tail;
        uint32_t total = t_tail;
        if (ch.meta->has_oxt) { const AtomRec o = pdb_oxt_rec(tt, ch); total += FCZ_PDB_LINE + atom_line_extra(o); }
        total += ter_line_len(pdb_last_rec(tt, ch));
        *scratch = total;
    }
    cx.sync();
    const uint32_t total = *scratch;
    cx.sync();
    return total;
}

struct alignas(16) V16 { uint32_t w[4]; };
// bytes from `src` to `dst` where both have the same 16-byte phase: 128-bit copies for the aligned interior
template <class Ctx>
FCZ_HD void copy_same_phase(Ctx& cx, char* dst, const char* src, uint32_t bytes) {
    const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    const uint32_t head = (16u - mis) & 15u;
    const uint32_t hb = head < bytes ? head : bytes;
    uint32_t body = 0;
    if (bytes > head) body = (bytes - head) & ~15u;
    for (uint32_t i = cx.tid; i < hb; i += cx.nthr) dst[i] = src[i];
    const V16* s4 = reinterpret_cast<const V16*>(src + head);
    V16* d4 = reinterpret_cast<V16*>(dst + head);
    for (uint32_t i = cx.tid; i < (body >> 4); i += cx.nthr) d4[i] = s4[i];
    for (uint32_t i = hb + body + cx.tid; i < bytes; i += cx.nthr) dst[i] = src[i];
}

// Emit the lines of residues [r_lo, r_hi) of a planned chain into its text (dst = first byte of the chain's
// text); the first unit also writes the TITLE lines, the last one the OXT and TER lines.  `stage` is a buffer of
// FCZ_PDB_STAGE_BYTES (16-byte aligned; shared memory on the device) through which uniform units -- every line
// 81 bytes -- leave as 128-bit copies; a unit with an over-long line writes straight to dst.
template <class Ctx>
FCZ_HD void pdb_emit_unit(Ctx& cx, const TextTables* tt, const PdbChain& ch, uint32_t r_lo, uint32_t r_hi, char* dst, char* stage,
                          uint32_t* us /* [2 * FCZ_PDB_UNIT_RES + 2] words visible to all threads */) {
    const uint32_t nr = r_hi - r_lo;
    uint32_t* ut = us + FCZ_PDB_UNIT_RES + 1u;
    for (uint32_t j = cx.tid; j <= nr; j += cx.nthr) { us[j] = ch.aoff[r_lo + j]; ut[j] = ch.toff[r_lo + j]; }
    cx.sync();
    const uint32_t a_lo = us[0], a_hi = us[nr];
    const uint32_t t_lo = ut[0], t_hi = ut[nr];
    const uint32_t n = a_hi - a_lo;
    const bool uniform = (t_hi - t_lo) == FCZ_PDB_LINE * n;
    char* g = dst + t_lo;
    char* s = stage + ((uintptr_t)g & 15u);
    for (uint32_t i = cx.tid; i < n; i += cx.nthr) {
        const uint32_t atom = a_lo + i;
        uint32_t j = 0, hi = nr;  // residue of this atom within the unit: the last j with us[j] <= atom
        while (hi - j > 1u) {
            const uint32_t mid = (j + hi) >> 1;
            if (us[mid] <= atom) j = mid; else hi = mid;
        }
        const uint32_t r = r_lo + j, k = atom - us[j];
        const AtomRec a = pdb_atom_rec(tt, ch, r, k, atom);
        if (uniform) {
            uint32_t w[21];
            atom_line_words(w, a);
            store_line81(s + FCZ_PDB_LINE * i, w);
        } else {
            uint32_t off = ut[j];
            for (uint32_t q = 0; q < k; q++) off += FCZ_PDB_LINE + atom_line_extra(pdb_atom_rec(tt, ch, r, q, us[j] + q));
            put_atom_line(dst + off, a);
        }
    }
    if (r_lo == 0u && cx.tid == 0) put_title_lines(dst, ch.title, ch.title_len);
    if (r_hi == ch.L && cx.tid == cx.nthr - 1) {
        char* p = dst + ch.toff[ch.L];
        if (ch.meta->has_oxt) p += put_atom_line(p, pdb_oxt_rec(tt, ch));
        put_ter_line(p, pdb_last_rec(tt, ch));
    }
    if (uniform) {
        cx.sync();
        cx.copy_out_same_phase(g, s, FCZ_PDB_LINE * n);  // device: one bulk async copy shared -> global; host model: copy_same_phase
    }
    cx.sync();
}

// ---- Foldcomp::extract (src/foldcomp.cpp:1260-1336)
// type 0: pLDDT (the continuised B-factor bytes) as `digits` characters per residue, comma separated when
// digits > 1; type 1: the one-letter sequence.
FCZ_HD uint32_t extract_plddt_stride(uint32_t digits) { return digits == 1u ? 1u : (digits == 2u ? 3u : (digits == 3u ? 5u : 6u)); }
FCZ_HD uint32_t extract_len(uint32_t L, int type, uint32_t digits) {
    if (type == 1) return L;
    if (L == 0u) return 0u;
    return digits == 1u ? L : L * extract_plddt_stride(digits) - 1u;  // no comma after the last residue
}
// the characters of residue i's pLDDT; returns how many (without the comma)
FCZ_HD uint32_t put_plddt(char* dst, float v, uint32_t digits, bool zero_to_one) {
    float clamped;
    char d1, d2;
    if (zero_to_one) {
        clamped = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        d1 = (char)((int)(clamped * 10.0f) % 10) + '0';
        d2 = (char)((int)(clamped * 100.0f) % 10) + '0';
    } else {
        clamped = v < 0.0f ? 0.0f : (v > 100.0f ? 100.0f : v);
        d1 = (char)(int)(clamped / 10.0f) + '0';  // 100.0 prints ':' as in the reference
        d2 = (char)((int)clamped % 10) + '0';
    }
    uint32_t n = 0;
    dst[n++] = d1;
    if (digits > 1u) dst[n++] = d2;
    if (digits >= 3u) {
        dst[n++] = '.';
        dst[n++] = (char)((int)(clamped * 10.0f) % 10) + '0';
    }
    if (digits == 4u) dst[n++] = (char)((int)(clamped * 100.0f) % 10) + '0';
    return n;
}

// One blob's extract output (thread per residue).  `y` is the blob's layout; dst receives extract_len() bytes.
template <class Ctx>
FCZ_HD void extract_chain(Ctx& cx, const TextTables* tt, const uint8_t* blob, const Layout& y, int type, uint32_t digits, char* dst) {
    const uint32_t L = y.L;
    if (type == 1) {
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) dst[r] = (char)tt->name1[blob[y.o_rec + 8u * r] >> 3];
        return;
    }
    const float tmin = get_f32(blob + y.o_temp), tcont = get_f32(blob + y.o_temp + 4u);
    // src/foldcomp.cpp:1290-1294: cont_f * (pow(2, 8) - 1) + min evaluated in double, compared as float
    const float maxval = (float)((double)tcont * 255.0 + (double)tmin);
    const bool z1 = maxval <= 1.0f && digits <= 2u;
    const uint32_t stride = extract_plddt_stride(digits);
    for (uint32_t i = cx.tid; i < L; i += cx.nthr) {
        const float v = continuize((unsigned)blob[y.o_temp + 8u + i], tmin, tcont);
        char* p = dst + i * stride;
        const uint32_t n = put_plddt(p, v, digits, z1);
        if (digits > 1u && i != L - 1u) p[n] = ',';
    }
}

}  // namespace fcz
#endif  // FCZ_TEXT_H
