/* include/fcz_engine.h -- C ABI of the B200 FCZ encode/decode engine.
 *
 * This is the drop-in boundary for Foldcomp's per-chain codec.  The
 * reference has no FFI for this path; its seam is the C++ class `Foldcomp`
 * (/root/reference/src/foldcomp.h:267-402) that the CLI lambdas (src/main.cpp:438-536 encode,
 * 612-689 decode) and the CPython module (foldcomp/foldcomp.cxx:197-220, 253-293) call once per
 * chain.  The entry points below replace, for a BATCH of chains at once:
 *
 *   fcz_encode_batch   <-  Foldcomp::compress()   src/foldcomp.cpp:562-606  (+ preprocess 450-559)
 *                          Foldcomp::writeStream() src/foldcomp.cpp:1038-1109
 *   fcz_decode_plan    <-  the header part of Foldcomp::read()  src/foldcomp.cpp:904-924
 *   fcz_decode_batch   <-  Foldcomp::read()        src/foldcomp.cpp:904-1036
 *                          Foldcomp::decompress()  src/foldcomp.cpp:779-902
 *   per-chain status   <-  read()'s int return (0 ok, -1 bad magic, src/foldcomp.cpp:911-915),
 *                          ValidityError (src/foldcomp.h:59-67), and the std::out_of_range that
 *                          escapes AAS.at() for residue names outside the table (src/sidechain.cpp:177)
 *
 * Chains are exchanged in a string-free canonical SoA layout: every residue carries a 5-bit type
 * code (src/utility.h:133-205) and its heavy atoms in the fixed slot order of the reference's
 * AminoAcid::atoms table (src/amino_acid.h:69-406; N, CA, C, O, CB, ...), FCZ_NATOMS[code] atoms
 * per residue, a missing atom given as (0,0,0) exactly as findFirstAtomCoords() returns it
 * (src/sidechain.cpp:140-147).  The host adapter (foldcomp_b200/csrc/foldcomp_gpu.h) converts
 * std::vector<AtomCoordinate>-style records to and from this layout.
 *
 * All pointers inside a batch live in ONE memory space, named by `mem`:
 *   FCZ_MEM_HOST    host memory (pageable or pinned); the call copies in/out and returns when the
 *                   results are on the host.
 *   FCZ_MEM_DEVICE  device memory of the engine's GPU; the call enqueues its kernels on the engine's
 *                   stream and returns without waiting for them (results are valid after the stream
 *                   syncs).  Calls that size something from the data wait for ONE small plan result
 *                   first: fcz_encode_batch (bytes per length tier, total size against bytes_cap),
 *                   the *_plan calls and fcz_extract_batch / fcz_unpack_angles_batch (their totals);
 *                   fcz_decode_batch, fcz_pdb_text_batch and fcz_parse_pdb_batch after their plan,
 *                   fcz_check_batch and fcz_backbone_angles_batch never wait.
 * Buffers are caller-owned; the engine owns its stream-ordered scratch.  One engine per GPU; calls
 * on one engine must be serialised by the caller, different engines are independent.
 */
#ifndef FCZ_ENGINE_H
#define FCZ_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCZ_MEM_HOST 0
#define FCZ_MEM_DEVICE 1

/* return / status codes */
#define FCZ_OK 0
#define FCZ_E_MAGIC (-1)         /* blob does not start with "FCMP"        (read() == -1)        */
#define FCZ_E_TRUNCATED (-2)     /* blob shorter than its header implies                          */
#define FCZ_E_RESIDUE (-3)       /* residue code outside the 20 amino acids + UNK (AAS.at throws) */
#define FCZ_E_LIMIT (-4)         /* nResidue/nAtom > 65535, nAnchor > 255 (src/foldcomp.h:118-131) or L < 2 */
#define FCZ_E_CAPACITY (-5)      /* an output buffer is too small                                 */
#define FCZ_E_CUDA (-6)          /* CUDA runtime error (see fcz_last_error)                       */
#define FCZ_E_ARG (-7)           /* bad argument                                                  */
/* PDB text (fcz_parse_pdb_*, fcz_encode_pdb_text_batch), per entry; the reference's parser flags (foldcomp/foldcomp.cxx:262-291) */
#define FCZ_E_PARSE_NOATOM (-11) /* no ATOM record ("No ATOM lines found", flag 1)                 */
#define FCZ_E_PARSE_CHAINS (-12) /* ATOM records of more than one chain ("Multiple chains", flag 2) */
#define FCZ_E_PARSE_RECORD (-13) /* an ATOM record too short for its columns (the reference throws std::out_of_range) */
#define FCZ_E_PARSE_NUMBER (-14) /* a numeric field that is not a plain fixed-point number (see fcz_parse_pdb_plan)   */
#define FCZ_E_PARSE_GAPS (-15)   /* the residue number of an N atom exceeds the previous N atom's by more than one, or the
                                    text does not start at an N atom: `foldcomp compress` cuts such a chain into fragments
                                    (identifyDiscontinousResInd, src/atom_coordinate.cpp:506-530; src/main.cpp:469-484)
                                    and encodes each on its own -- the entry is not ONE chain (the CPython compress() does
                                    not cut, and stores anchors picked by residue NUMBER, src/foldcomp.cpp:756-760)      */

typedef struct fcz_engine fcz_engine;

typedef struct fcz_opts {
    int32_t anchor_threshold;   /* Foldcomp::anchorThreshold, CLI -b/--break; default 25 (src/foldcomp.h:56) */
    int32_t use_alt_atom_order; /* Foldcomp::useAltAtomOrder, CLI -a/--alt (decode only)          */
    void* stream;               /* cudaStream_t to enqueue on; NULL = engine-owned stream         */
    int32_t terminate_blobs;    /* encode: one NUL byte after every blob (counted in blob_off), so that `bytes`
                                   is a foldcomp-db data slab as it stands: entries NUL-terminated like
                                   `decompress --db` and test/example_db (src/main.cpp:656-665) */
} fcz_opts;

/* Per-chain scalars that the reference keeps in Foldcomp members / CompressedFileHeader
 * (src/foldcomp.h:118-136). */
typedef struct fcz_chain_meta {
    uint16_t n_atom;      /* header nAtom: atom count of the ORIGINAL input incl. OXT            */
    uint16_t idx_residue; /* residue number of the first residue                                 */
    uint16_t idx_atom;    /* serial of the first atom                                            */
    uint8_t chain;        /* chain id character                                                  */
    uint8_t has_oxt;      /* 1 when the input ended with an OXT atom (src/foldcomp.cpp:473-481)  */
    float oxt[3];         /* its coordinates, zeros otherwise                                    */
} fcz_chain_meta;

/* A batch of chains in canonical slot order.  Input of encode, output of decode. */
typedef struct fcz_chain_batch {
    uint32_t n_chains;
    int32_t mem;          /* FCZ_MEM_HOST | FCZ_MEM_DEVICE for every pointer below               */
    uint32_t* res_off;    /* [n_chains+1] first residue of each chain                            */
    uint64_t* atom_off;   /* [n_chains+1] first atom of each chain (sum of FCZ_NATOMS)           */
    uint32_t* title_off;  /* [n_chains+1] first title byte of each chain                         */
    uint8_t* res_type;    /* [n_res]   5-bit residue codes                                       */
    float* bfactor;       /* [n_res]   B-factor / pLDDT of the residue's CA                      */
    float* xyz;           /* [3*n_atoms] x,y,z per atom, residue-major, slot order               */
    char* titles;         /* [n_title_bytes] concatenated titles, no terminators                 */
    fcz_chain_meta* meta; /* [n_chains]                                                          */
    int32_t* status;      /* [n_chains] per-chain status written by decode (may be NULL on encode input) */
    /* capacities (elements) of res_type/bfactor, xyz/3 and titles when used as decode output     */
    uint64_t res_cap, atom_cap, title_cap;
} fcz_chain_batch;

/* A batch of FCZ blobs, tightly concatenated.  Output of encode, input of decode. */
typedef struct fcz_blob_batch {
    uint32_t n_chains;
    int32_t mem;
    uint64_t* blob_off; /* [n_chains+1] byte offset of each blob in `bytes`                      */
    uint8_t* bytes;
    int32_t* status;    /* [n_chains] per-chain status written by encode (may be NULL on decode input) */
    uint64_t bytes_cap; /* capacity of `bytes` when used as encode output                        */
} fcz_blob_batch;

/* totals returned by the planning calls */
typedef struct fcz_sizes {
    uint64_t n_res, n_atoms, n_title_bytes, n_blob_bytes;
} fcz_sizes;

fcz_engine* fcz_engine_create(int device, const fcz_opts* opts);
void fcz_engine_destroy(fcz_engine* e);
int fcz_engine_set_opts(fcz_engine* e, const fcz_opts* opts);

/* Upper bound of the encoded size of a batch, from totals only (host arithmetic, no GPU work):
 * sum over chains of 97 + 40*nAnchor + lenTitle + 8*L + (atoms - 3*L) + L  (SURVEY.md Appendix A), plus one
 * byte per chain for opts.terminate_blobs. */
uint64_t fcz_encode_bound(uint64_t n_chains, uint64_t n_res, uint64_t n_atoms,
                          uint64_t n_title_bytes, int32_t anchor_threshold);

/* Encode every chain of `in` into `out`.  Fills out->blob_off[0..n], out->bytes and out->status.
 * Blobs are byte-identical to the reference's writeStream() output except that the four padding
 * bytes of CompressedFileHeader (file offsets 14,15,22,23), which the reference leaves
 * uninitialised, are written as zero.  A chain with a non-zero status gets an empty blob. */
int fcz_encode_batch(fcz_engine* e, const fcz_chain_batch* in, fcz_blob_batch* out);

/* Read the blob headers and fill out->res_off, out->atom_off, out->title_off (n_chains+1 each)
 * and out->status.  `totals` (host memory) receives the sizes the caller must provide to
 * fcz_decode_batch; this call synchronises the engine's stream. */
int fcz_decode_plan(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals);

/* Decode every blob of `in` into `out` (whose offset arrays were filled by fcz_decode_plan and
 * whose data arrays have at least the planned capacities).  Atom order inside a residue is the
 * canonical slot order, or FCZ_ALT order when opts.use_alt_atom_order is set. */
int fcz_decode_batch(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out);

/* ---- text ------------------------------------------------------------------
 * A batch of texts, tightly concatenated: PDB text per chain, or extract output per blob. */
typedef struct fcz_text_batch {
    uint32_t n_chains;
    int32_t mem;
    uint64_t* text_off;  /* [n_chains+1] byte offset of each chain's text in `bytes`              */
    char* bytes;
    uint64_t bytes_cap;  /* capacity of `bytes`                                                  */
    int32_t* status;     /* [n_chains] per-chain status, written by fcz_decode_to_pdb_plan (may be NULL) */
} fcz_text_batch;

/* PDB text of every chain of a decoded batch, byte-identical to the reference's
 * writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291; fast_ftoa 186-218) on the atoms that
 * Foldcomp::decompress returns (src/foldcomp.cpp:779-902): TITLE lines, one ATOM line per atom (serials from
 * meta.idx_atom, residue numbers from meta.idx_residue, the OXT record with residue number = nResidue as in
 * src/foldcomp.cpp:958-961), TER.  Atom names follow opts.use_alt_atom_order, i.e. the order the decode call
 * produced.  fcz_pdb_text_plan fills out->text_off and returns the total size (it synchronises the engine's
 * stream); fcz_pdb_text_batch must follow it on the same batch with out->bytes of at least that capacity.
 * For FCZ_MEM_DEVICE batches in->res_cap must be >= the batch's residue count. */
int fcz_pdb_text_plan(fcz_engine* e, const fcz_chain_batch* in, fcz_text_batch* out, uint64_t* total_bytes);
int fcz_pdb_text_batch(fcz_engine* e, const fcz_chain_batch* in, fcz_text_batch* out);

/* Decode and format in one go, host blobs in, host text out: what `foldcomp decompress` does per entry
 * (src/main.cpp:612-689: Foldcomp::read + decompress + writeAtomCoordinatesToPDB).  The blobs go up once, the
 * decoded coordinates stay on the GPU, only the text comes back (in slabs, overlapped with the emit kernel).
 * Same plan/batch protocol as above; a blob that fails to decode yields an empty text and a status. */
int fcz_decode_to_pdb_plan(fcz_engine* e, const fcz_blob_batch* in, fcz_text_batch* out, uint64_t* total_bytes);
int fcz_decode_to_pdb_batch(fcz_engine* e, const fcz_blob_batch* in, fcz_text_batch* out);

/* Foldcomp::extract (src/foldcomp.cpp:1260-1336) for every blob: type 0 = pLDDT with `digits` (1..4) characters
 * per residue, comma separated when digits > 1; type 1 = one-letter amino-acid sequence.  A blob that fails the
 * header check yields an empty text.  *total_bytes receives the size needed; FCZ_E_CAPACITY if out is smaller. */
int fcz_extract_batch(fcz_engine* e, const fcz_blob_batch* in, int32_t type, int32_t digits, fcz_text_batch* out,
                      uint64_t* total_bytes);

/* Foldcomp::read (src/foldcomp.cpp:904-1036) followed by Foldcomp::checkValidity (src/foldcomp.cpp:1492-1532; classes
 * ValidityError src/foldcomp.h:59-67) for every blob: what `foldcomp check` (src/main.cpp:910-928) and `decompress
 * --check` (630-636) run per entry.
 *   read_status[c]  0, or FCZ_E_MAGIC where read() returns -1; FCZ_E_TRUNCATED where the blob ends before the sections its
 *                   header announces (the reference reads past the end of its stream there and checks whatever its
 *                   buffers held; here that is an error of its own)
 *   validity[c]     the ValidityError class, FCZ_V_*; for a truncated blob the COUNT_MISMATCH class of the first section
 *                   that falls short; FCZ_V_SUCCESS for a blob that fails the magic check (nothing was read)
 * Order of the checks as in the reference: backbone (every record has phi = psi = omega = 0), side chain (every byte 0 --
 * and, std::all_of over an empty range being true, a chain WITHOUT side-chain torsions), B-factors (every byte 0).
 * Both arrays live in the memory space of `in`, n_chains entries each; either may be NULL. */
#define FCZ_V_SUCCESS 0
#define FCZ_V_BACKBONE_COUNT_MISMATCH 1
#define FCZ_V_SIDECHAIN_COUNT_MISMATCH 2
#define FCZ_V_TEMP_FACTOR_COUNT_MISMATCH 3
#define FCZ_V_EMPTY_BACKBONE_ANGLE 4
#define FCZ_V_EMPTY_SIDECHAIN_ANGLE 5
#define FCZ_V_EMPTY_TEMP_FACTOR 6
int fcz_check_batch(fcz_engine* e, const fcz_blob_batch* in, int32_t* read_status, int32_t* validity);

/* ---- PDB text in ----------------------------------------------------------------------------
 * The fixed-column ATOM parser of the reference's CPython compress() (foldcomp/foldcomp.cxx:253-293: atom 12-15, residue
 * 17-19, chain 21, serial 6-10, residue number 22-25, x y z 30-53, B-factor 60-65; removeAlternativePosition
 * src/atom_coordinate.cpp:362-370) followed by the encoder's by-name bookkeeping (residue split src/atom_coordinate.cpp:304-328,
 * first atom of every table name src/sidechain.cpp:140-147, CA B-factor src/foldcomp.cpp:543-547, OXT 473-481), for a
 * batch of single-chain PDB texts on the GPU: `in` = texts tightly concatenated (text_off[n+1], bytes), DEVICE memory;
 * `out` = the canonical chain batch in DEVICE memory.  fcz_parse_pdb_plan fills out->res_off, out->atom_off, out->status
 * (0 or FCZ_E_PARSE_*: such an entry has no residues) and `totals` (it synchronises the engine's stream);
 * fcz_parse_pdb_batch must follow on the same batch and fills res_type, bfactor, xyz, meta.  Titles are the caller's
 * (entry names): title_off / titles are not touched.  Numeric fields are converted exactly like the reference's
 * std::stof when they are plain fixed-point numbers of at most nine digits (what every PDB writer emits); any other shape
 * (exponent, hex float, nan) makes the entry FCZ_E_PARSE_NUMBER. */
int fcz_parse_pdb_plan(fcz_engine* e, const fcz_text_batch* in, fcz_chain_batch* out, fcz_sizes* totals);
int fcz_parse_pdb_batch(fcz_engine* e, const fcz_text_batch* in, fcz_chain_batch* out);

/* PDB texts in HOST memory -> FCZ blobs in HOST memory in one call: what `foldcomp compress` does per entry
 * (src/main.cpp:438-536) for a batch.  The text goes up once, parsing and encoding stay on the GPU, only the blobs come
 * back.  title_off[n+1] / titles: the entries' titles (host).  out->blob_off, out->bytes (capacity out->bytes_cap),
 * out->status (parser or encoder status per entry; a failed entry gets an empty blob).  *total_bytes receives the size
 * needed; FCZ_E_CAPACITY when out->bytes is smaller. */
int fcz_encode_pdb_text_batch(fcz_engine* e, const fcz_text_batch* in, const uint32_t* title_off, const char* titles,
                              fcz_blob_batch* out, uint64_t* total_bytes);

/* Continuised backbone angles of every residue record: six floats per residue, phi, psi, omega, N-CA-C, CA-C-N, C-N-CA
 * -- decompressBackboneChain (src/foldcomp.cpp:122-153), i.e. what Foldcomp::decompress leaves in its phi / psi / omega /
 * *_angle members (783-804) and the CPython get_data() returns (foldcomp/foldcomp.cxx:497-560).  Fills res_off[0..n]
 * (residues before each blob; a blob that fails the header check counts 0) and, when `angles` is not NULL and
 * res_cap is large enough, angles[6 * residues]; *total_res receives the residue total either way. */
int fcz_unpack_angles_batch(fcz_engine* e, const fcz_blob_batch* in, uint64_t* res_off, float* angles, uint64_t res_cap,
                            uint64_t* total_res);

/* Backbone angles BEFORE quantisation, six floats per residue r of every chain: the torsions over backbone atoms 3r+j ..
 * 3r+j+3, j = 0..2 (psi_r, omega_r, phi_r+1; zero for a chain's last residue), then the bond angles at backbone atoms 3r,
 * 3r+1, 3r+2 (zero at a chain's first and last atom) -- what Foldcomp::preprocess leaves in backboneTorsionAngles
 * (getTorsionFromXYZ, src/torsion_angle.cpp:46-96) and backboneBondAngles (Nerf::getBondAngles, src/nerf.cpp:495-508) at
 * src/foldcomp.cpp:484-496, bit for bit; the CPython get_data(pdb_text) returns them (foldcomp/foldcomp.cxx:633-671).
 * `angles` [6 * n_res] lives in the memory space of `in`; a residue code outside the table reads as UNK has no atoms:
 * chains must have passed an encode (or come from the parser).  Device batches need in->res_cap >= n_res. */
int fcz_backbone_angles_batch(fcz_engine* e, const fcz_chain_batch* in, float* angles);

/* Page-locked host memory for the buffers of FCZ_MEM_HOST batches (cudaHostAlloc / cudaFreeHost): copies from and to such
 * memory run at the full link rate and asynchronously; pageable memory works everywhere too, at roughly half the rate.
 * For host code that does not link the CUDA runtime itself (foldcomp_b200/csrc/fcz_db.cpp, bindings in other languages). */
void* fcz_host_alloc(size_t bytes);
void fcz_host_free(void* p);

/* Block until everything enqueued on the engine's stream has finished. */
int fcz_engine_sync(fcz_engine* e);

/* Number of kernels this engine has launched since creation (bench.py's gpu_launches). */
uint64_t fcz_engine_launch_count(const fcz_engine* e);

/* Optional timing with CUDA events recorded around every launch, on the stream the kernel is launched on.
 * fcz_engine_get_profile synchronises the engine's stream, returns the accumulated device times since the
 * last call and resets them.  bench.py uses it for the roofline of the dominant kernel.
 *   encode_kernel_ms / decode_kernel_ms   whole hot-path spans: every tier's launches, fork to join
 *   kernel_ms[k] / kernel_launches[k]     the same per kind k (FCZ_PROF_*): the two spans and each kernel alone */
#define FCZ_PROF_ENCODE 0        /* span: all k_encode launches of one fcz_encode_batch                        */
#define FCZ_PROF_DECODE 1        /* span: front + stitch + back of every tier of one fcz_decode_batch          */
#define FCZ_PROF_K_ENCODE 2      /* one k_encode launch                                                        */
#define FCZ_PROF_K_DEC_FRONT 3   /* one k_dec_front launch (unpack + NeRF passes)                              */
#define FCZ_PROF_K_DEC_STITCH 4  /* one k_dec_stitch_t launch (serial walk over anchor segments)               */
#define FCZ_PROF_K_DEC_BACK 5    /* one k_dec_back launch (blend + side chains + copy-out)                     */
#define FCZ_PROF_K_PDB_PLAN 6    /* one k_pdb_plan launch                                                      */
#define FCZ_PROF_K_PDB_EMIT 7    /* one k_pdb_emit launch                                                      */
#define FCZ_PROF_KINDS 8
typedef struct fcz_profile {
    double encode_kernel_ms, decode_kernel_ms;
    uint64_t encode_launches, decode_launches;
    double kernel_ms[FCZ_PROF_KINDS];
    uint64_t kernel_launches[FCZ_PROF_KINDS];
} fcz_profile;
int fcz_engine_set_profiling(fcz_engine* e, int enabled);
int fcz_engine_get_profile(fcz_engine* e, fcz_profile* out);

const char* fcz_strerror(int code);
const char* fcz_last_error(const fcz_engine* e);

/* Residue tables (foldcomp_b200/csrc/fcz_tables.h) for host code that cannot include the header. */
int fcz_type_natoms(int code);
const char* fcz_type_name3(int code);
const char* fcz_type_atom_name(int code, int slot);
int fcz_type_alt_slot(int code, int pos);
int fcz_type_pred(int code, int slot, int which);
float fcz_type_bond_length(int code, int slot);
float fcz_type_bond_angle(int code, int slot);

#ifdef __cplusplus
}
#endif
#endif /* FCZ_ENGINE_H */
