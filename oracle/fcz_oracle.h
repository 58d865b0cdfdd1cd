/* oracle/fcz_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, scalar, CPU restatement of the reference's FCZ encode/decode algorithm
 * (steineggerlab/foldcomp @ 30496eb), used ONLY as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
 * (foldcomp_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against the unmodified
 * reference compiled in this container (oracle/_ref/libfoldcomp_ref.so): FCZ bytes bit-identical
 * (modulo the four uninitialised header padding bytes), decoded coordinates bit-identical, and
 * repeats it against committed fixtures (tests/golden/golden.npz) generated from the reference by
 * tests/golden/make_golden.py.  The text functions (PDB writer, extract, continuised angles) are
 * pinned the same way by tests/test_text.py (tests/golden/text_golden.npz, make_text_golden.py) and
 * tests/test_db_host.py (the reference's own CPython module, oracle/_ref/pyref).
 */
#ifndef FCZ_ORACLE_H
#define FCZ_ORACLE_H
#include "../include/fcz_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Encode one chain (canonical slot layout).  Returns blob size, or a negative FCZ_E_* code.
 * out may be NULL to query the size. */
int64_t fcz_oracle_encode_chain(const uint8_t* res_type, uint32_t L, const float* xyz,
                                const float* bfactor, const fcz_chain_meta* meta, const char* title,
                                uint32_t title_len, int32_t anchor_threshold, uint8_t* out,
                                uint64_t cap);

/* Header peek: residues, decoded atoms (sum of table atoms + OXT), title length. */
int fcz_oracle_peek(const uint8_t* blob, uint64_t len, uint32_t* L, uint64_t* n_atoms,
                    uint32_t* title_len);

/* Decode one blob.  xyz gets n_atoms*3 floats (OXT last if present), res_type/bfactor L entries. */
int fcz_oracle_decode_chain(const uint8_t* blob, uint64_t len, int use_alt, uint8_t* res_type,
                            float* bfactor, float* xyz, fcz_chain_meta* meta, char* title);

/* Batch versions over host-memory batches (same structs as the engine ABI); OpenMP over chains. */
int fcz_oracle_encode_batch(const fcz_chain_batch* in, fcz_blob_batch* out, int32_t anchor_threshold,
                            int n_threads);
int fcz_oracle_decode_plan(const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals);
int fcz_oracle_decode_batch(const fcz_blob_batch* in, fcz_chain_batch* out, int use_alt, int n_threads);

/* PDB text of one chain in the decoder's output layout: the reference's writeAtomCoordinatesToPDB
 * (src/atom_coordinate.cpp:220-291) with fast_ftoa (186-218).  Returns the text length. */
int64_t fcz_oracle_format_pdb(const uint8_t* res_type, uint32_t L, const float* xyz, const float* bfactor,
                              const fcz_chain_meta* meta, const char* title, uint32_t title_len, int use_alt,
                              char* out, uint64_t cap);

/* Foldcomp::extract (src/foldcomp.cpp:1260-1336): type 0 = pLDDT with `digits` 1..4, type 1 = sequence. */
int64_t fcz_oracle_extract(const uint8_t* blob, uint64_t len, int type, int digits, char* out, uint64_t cap);

/* Continuised angles of every residue record, 6 floats each (phi, psi, omega, N-CA-C, CA-C-N, C-N-CA):
 * decompressBackboneChain, src/foldcomp.cpp:122-153.  Returns the residue count or a negative code. */
int64_t fcz_oracle_unpack_angles(const uint8_t* blob, uint64_t len, float* out);
/* backbone angles before quantisation, out[6 * L] (see fcz_oracle.c) */
int fcz_oracle_backbone_angles(const uint8_t* res_type, uint32_t L, const float* xyz, float* out);

/* Foldcomp::read + checkValidity (src/foldcomp.cpp:904-1036, 1492-1532): returns the read status, *validity = the
 * ValidityError class (src/foldcomp.h:59-67). */
int fcz_oracle_check(const uint8_t* blob, uint64_t len, int* validity);

#ifdef __cplusplus
}
#endif
#endif
