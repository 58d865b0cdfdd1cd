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

#include "fcz_format.h"
#include "fcz_tables.h"

namespace fcz {

#define FCZ_PDB_LINE 81u  // bytes of a uniform ATOM line
#define FCZ_PDB_TER 27u   // bytes of a uniform TER line

// Names as the emitter needs them: residue names and atom names packed little-endian into 32-bit words, atom
// names already left-justified in 3 columns (std::setw(3) << std::left, src/atom_coordinate.cpp:253-255).
struct TextTables {
    uint32_t name3[FCZ_NUM_CODES];                 // 'A' | 'L' << 8 | 'A' << 16
    uint32_t atom[FCZ_NUM_CODES][FCZ_MAX_ATOMS];   // "CA " etc; byte 0 is also the element column (atom[0])
    uint8_t natoms[FCZ_NUM_CODES];
    uint8_t alt[FCZ_NUM_CODES][FCZ_MAX_ATOMS];
    uint8_t name1[FCZ_NUM_CODES];
};

inline void build_text_tables(TextTables* t) {
    for (int c = 0; c < FCZ_NUM_CODES; c++) {
        t->name3[c] = (uint32_t)(uint8_t)FCZ_NAME3[c][0] | (uint32_t)(uint8_t)FCZ_NAME3[c][1] << 8 | (uint32_t)(uint8_t)FCZ_NAME3[c][2] << 16;
        t->natoms[c] = FCZ_NATOMS[c];
        t->name1[c] = (uint8_t)FCZ_NAME1[c];
        for (int k = 0; k < FCZ_MAX_ATOMS; k++) {
            uint32_t w = 0;
            bool end = false;
            for (int j = 0; j < 3; j++) {
                char ch = end ? ' ' : FCZ_ATOM_NAME[c][k][j];
                if (ch == 0) { end = true; ch = ' '; }
                w |= (uint32_t)(uint8_t)ch << (8 * j);
            }
            t->atom[c][k] = w;
            t->alt[c][k] = FCZ_ALT[c][k];
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

}  // namespace fcz
#endif  // FCZ_TEXT_H
