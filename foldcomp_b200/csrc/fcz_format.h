// foldcomp_b200/csrc/fcz_format.h -- FCZ wire format: section layout and byte helpers (host + device).
// Reference: Foldcomp::writeStream src/foldcomp.cpp:1038-1109, Foldcomp::read 904-1036,
// CompressedFileHeader src/foldcomp.h:118-136, getSize src/foldcomp.cpp:1190-1214.
#ifndef FCZ_FORMAT_H
#define FCZ_FORMAT_H

#include "fcz_math.h"
#include "fcz_tables.h"

namespace fcz {

// file offsets inside magic + CompressedFileHeader (little-endian host struct written raw)
enum {
    OFF_MAGIC = 0,
    OFF_NRES = 4,      // u16 nResidue
    OFF_NATOM = 6,     // u16 nAtom
    OFF_IDXRES = 8,    // u16 idxResidue
    OFF_IDXATOM = 10,  // u16 idxAtom
    OFF_NANCHOR = 12,  // u8
    OFF_CHAIN = 13,    // char   (+2 padding bytes 14,15: zero here, uninitialised in the reference)
    OFF_NSC = 16,      // u32 nSideChainTorsion
    OFF_FIRSTRES = 20, // char
    OFF_LASTRES = 21,  // char   (+2 padding bytes 22,23)
    OFF_LENTITLE = 24, // u32
    OFF_MINS = 28,     // f32[6]  phi, psi, omega, n_ca_c, ca_c_n, c_n_ca (src/foldcomp.cpp:1354-1365)
    OFF_CONTFS = 52,   // f32[6]
    HDR_BYTES = 76
};

// rows of the device-side residue tables (5-bit code space) and the reference's mapping of unknown codes to UNK
#define FCZ_CODE_ROWS 32
FCZ_HD unsigned norm_code(unsigned code) { return code < (unsigned)FCZ_NUM_CODES ? code : (unsigned)FCZ_CODE_UNK; }
// Table atom count of a 5-bit residue code (rows 24..31 = UNK, like Tables::natoms) from two 64-bit constants, four bits per
// code: pure arithmetic where a table look-up would be a dependent load (build_tables checks it against FCZ_NATOMS).
FCZ_HD uint32_t natoms_packed(unsigned code5) {
    const unsigned long long k = (code5 & 16u) ? 0x3333333330007ce7ull : 0x67b8988a499688b5ull;
    return (uint32_t)(k >> (4u * (code5 & 15u))) & 15u;
}

// header order of the six backbone arrays
enum { A_PHI = 0, A_PSI = 1, A_OMEGA = 2, A_NCAC = 3, A_CACN = 4, A_CNCA = 5 };

// number of bins: NUM_BITS_* of src/foldcomp.h:44-49, Discretizer(values, pow(2,bits)-1)
FCZ_HD unsigned n_bins(int k) { return k < 2 ? 4095u : (k == 2 ? 2047u : 255u); }

// ideal backbone bond lengths, src/foldcomp.h:51-54 (double literals narrowed to float at the call)
#define FCZ_N_TO_CA ((float)1.4581)
#define FCZ_CA_TO_C ((float)1.5281)
#define FCZ_C_TO_N ((float)1.3311)
#define FCZ_PRO_N_TO_CA ((float)1.353)

// reference: _getAnchorNum/_setAnchor, src/foldcomp.cpp:739-761
FCZ_HD int anchor_count(uint32_t L, int b) { return (int)L / b + 2; }
FCZ_HD int anchor_index(uint32_t L, int n_all, int i) {
    int interval = (int)L / (n_all - 1);
    return (i == n_all - 1) ? (int)L - 1 : i * interval;
}

// section offsets of one blob
struct Layout {
    uint32_t L, n_sc, title_len, n_anchor;
    uint32_t o_aidx, o_title, o_anchor, o_oxt, o_rec, o_sc, o_temp, size;
};
FCZ_HD Layout make_layout(uint32_t L, uint32_t n_sc, uint32_t title_len, uint32_t n_anchor) {
    Layout y;
    y.L = L; y.n_sc = n_sc; y.title_len = title_len; y.n_anchor = n_anchor;
    y.o_aidx = HDR_BYTES;
    y.o_title = y.o_aidx + 4u * n_anchor;
    y.o_anchor = y.o_title + title_len;
    y.o_oxt = y.o_anchor + 36u * n_anchor;
    y.o_rec = y.o_oxt + 13u;
    y.o_sc = y.o_rec + 8u * L;
    y.o_temp = y.o_sc + n_sc;
    y.size = y.o_temp + 8u + L;
    return y;
}

// unaligned little-endian accessors (blobs are tightly packed, so nothing is aligned)
FCZ_HD void put_u16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
FCZ_HD void put_u32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
FCZ_HD uint32_t get_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
FCZ_HD uint32_t get_u32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
FCZ_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
#endif
}
FCZ_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
#endif
}
FCZ_HD void put_f32(uint8_t* p, float v) { put_u32(p, f2u(v)); }
FCZ_HD float get_f32(const uint8_t* p) { return u2f(get_u32(p)); }

// reference: convertBackboneChainToBytes, src/foldcomp.cpp:33-52 (bit-fields of struct
// BackboneChain, src/foldcomp.h:71-81, truncate each value to its width)
FCZ_HD void pack_record(uint8_t* p, unsigned res, unsigned phi, unsigned psi, unsigned omg, unsigned nca,
                        unsigned cac, unsigned cnc) {
    res &= 0x1F; omg &= 0x7FF; psi &= 0xFFF; phi &= 0xFFF;
    p[0] = (uint8_t)((res << 3) | (omg >> 8));
    p[1] = (uint8_t)(omg & 0xFF);
    p[2] = (uint8_t)(psi >> 4);
    p[3] = (uint8_t)(((psi & 0xF) << 4) | (phi >> 8));
    p[4] = (uint8_t)(phi & 0xFF);
    p[5] = (uint8_t)cac;
    p[6] = (uint8_t)cnc;
    p[7] = (uint8_t)nca;
}
// reference: convertBytesToBackboneChain, src/foldcomp.cpp:60-77
struct Record {
    unsigned res, phi, psi, omg, nca, cac, cnc;
};
FCZ_HD Record unpack_record(const uint8_t* b) {
    Record r;
    r.res = b[0] >> 3;
    r.omg = ((b[0] & 7u) << 8) | b[1];
    r.psi = ((unsigned)b[2] << 4) | (b[3] >> 4);
    r.phi = ((b[3] & 0xFu) << 8) | b[4];
    r.cac = b[5]; r.cnc = b[6]; r.nca = b[7];
    return r;
}

}  // namespace fcz
#endif  // FCZ_FORMAT_H
