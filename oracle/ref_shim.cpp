// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C adapter around the UNMODIFIED reference implementation
// (/root/reference/src/*.cpp, compiled where they lie by oracle/Makefile into
// oracle/_ref/libfoldcomp_ref.so).  It converts the canonical SoA chain layout
// used by this repo (include/fcz_engine.h) to/from the reference's
// std::vector<AtomCoordinate> and then calls the reference's own
//   Foldcomp::compress  (src/foldcomp.cpp:562)   + writeStream (src/foldcomp.cpp:1038)
//   Foldcomp::read      (src/foldcomp.cpp:904)   + decompress  (src/foldcomp.cpp:779)
// Nothing here re-implements any arithmetic; atom names per (residue type, slot)
// come from the reference's own AminoAcid::AminoAcids() table (src/amino_acid.h:69).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
#include "amino_acid.h"
#include "foldcomp.h"
#include "atom_coordinate.h"
#include "utility.h"

#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const std::map<std::string, AminoAcid>& aas() {
    static const std::map<std::string, AminoAcid> m = AminoAcid::AminoAcids();
    return m;
}

struct TypeInfo {
    std::string name3;
    std::vector<std::string> atoms;  // canonical slot order (N, CA, C, O, CB, ...)
    std::vector<int> alt_perm;       // alt_perm[j] = canonical slot of j-th atom in -a order
};

const TypeInfo& type_info(int code) {
    // thread-safe one-time initialisation (C++11 magic static): ref_roundtrip_batch calls this from OpenMP threads
    static const std::vector<TypeInfo> tab = [] {
        std::vector<TypeInfo> v(24);
        for (int c = 0; c < 24; c++) {
            TypeInfo& t = v[c];
            t.name3 = convertIntToThreeLetterCode((unsigned)c);
            auto it = aas().find(t.name3);
            if (it != aas().end() && !it->second.atoms.empty()) {
                t.atoms = it->second.atoms;
            } else {
                t.atoms = {"N", "CA", "C"};
            }
        }
        return v;
    }();
    return tab[code];
}

void build_atoms(const uint8_t* res_type, int L, const float* xyz, const float* bfac,
                 int has_oxt, const float* oxt, int idx_res, int idx_atom, char chain,
                 std::vector<AtomCoordinate>& atoms) {
    std::string ch(1, chain);
    int ai = idx_atom;
    size_t a = 0;
    for (int r = 0; r < L; r++) {
        const TypeInfo& t = type_info(res_type[r]);
        for (size_t k = 0; k < t.atoms.size(); k++, a++) {
            atoms.emplace_back(t.atoms[k], t.name3, ch, ai++, idx_res + r,
                               xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2], 1.0f, bfac[r]);
        }
    }
    if (has_oxt) {
        const TypeInfo& t = type_info(res_type[L - 1]);
        atoms.emplace_back("OXT", t.name3, ch, ai++, idx_res + L - 1, oxt[0], oxt[1], oxt[2], 1.0f,
                           bfac[L - 1]);
    }
}

int compress_one(const uint8_t* res_type, int L, const float* xyz, const float* bfac, int has_oxt,
                 const float* oxt, int idx_res, int idx_atom, char chain, const char* title,
                 int title_len, int anchor_threshold, std::string& out) {
    std::vector<AtomCoordinate> atoms;
    build_atoms(res_type, L, xyz, bfac, has_oxt, oxt, idx_res, idx_atom, chain, atoms);
    Foldcomp comp;
    comp.strTitle = std::string(title, title + title_len);
    comp.anchorThreshold = anchor_threshold;
    tcb::span<AtomCoordinate> sp(atoms.data(), atoms.size());
    comp.compress(sp);
    std::ostringstream oss;
    comp.writeStream(oss);
    out = oss.str();
    return 0;
}

int decompress_one(const uint8_t* fcz, size_t len, int use_alt, std::vector<AtomCoordinate>& atoms,
                   Foldcomp& comp) {
    std::istringstream iss(std::string((const char*)fcz, len));
    comp.useAltAtomOrder = use_alt != 0;
    int rc = comp.read(iss);
    if (rc != 0) return rc;
    return comp.decompress(atoms);
}

}  // namespace

extern "C" {

// Number of canonical atom slots for a 5-bit residue code, and the slot names.
int ref_type_natoms(int code) { return (int)type_info(code).atoms.size(); }
const char* ref_type_atom_name(int code, int slot) { return type_info(code).atoms[slot].c_str(); }
const char* ref_type_name3(int code) { return type_info(code).name3.c_str(); }

// Compress one chain given in canonical slot order.  Returns 0 and the blob length in *out_len
// (-1 if cap is too small; *out_len is then the needed size).
int ref_compress(const uint8_t* res_type, int L, const float* xyz, const float* bfac, int has_oxt,
                 const float* oxt, int idx_res, int idx_atom, char chain, const char* title,
                 int title_len, int anchor_threshold, uint8_t* out, size_t cap, size_t* out_len) {
    std::string blob;
    try {
        compress_one(res_type, L, xyz, bfac, has_oxt, oxt, idx_res, idx_atom, chain, title,
                     title_len, anchor_threshold, blob);
    } catch (const std::exception&) {
        return -3;
    }
    *out_len = blob.size();
    if (blob.size() > cap) return -1;
    memcpy(out, blob.data(), blob.size());
    return 0;
}

// Decompress one blob.  xyz_out receives atoms in the reference's output order (canonical slot
// order, or the -a order when use_alt), OXT last when present.  bfac_out / res_type_out are per
// residue.  Returns the reference's return code (0 ok, -1 bad magic, -2).
int ref_decompress(const uint8_t* fcz, size_t len, int use_alt, float* xyz_out, size_t cap_atoms,
                   int* n_atoms, float* bfac_out, uint8_t* res_type_out, size_t cap_res, int* n_res,
                   int* has_oxt) {
    std::vector<AtomCoordinate> atoms;
    Foldcomp comp;
    int rc;
    try {
        rc = decompress_one(fcz, len, use_alt, atoms, comp);
    } catch (const std::exception&) {
        return -3;
    }
    if (rc != 0) return rc;
    *n_atoms = (int)atoms.size();
    *n_res = comp.nResidue;
    *has_oxt = comp.hasOXT;
    if (atoms.size() > cap_atoms || (size_t)comp.nResidue > cap_res) return -4;
    for (size_t i = 0; i < atoms.size(); i++) {
        xyz_out[3 * i] = atoms[i].coordinate.x;
        xyz_out[3 * i + 1] = atoms[i].coordinate.y;
        xyz_out[3 * i + 2] = atoms[i].coordinate.z;
    }
    for (int r = 0; r < comp.nResidue; r++) {
        bfac_out[r] = comp.tempFactors[r];
        res_type_out[r] = (uint8_t)comp.compressedBackBone[r].residue;
    }
    return 0;
}

// Batch round trip on host threads: compress + writeStream + read + decompress per chain, the
// "core-only" CPU baseline of BASELINE.md section 4.1.  Chains are given as in fcz_chain_batch
// (offset arrays); titles are "syn_%07d" of the chain index.  Writes nothing back except an
// xor-checksum (so the work cannot be optimised away) and the total blob bytes.
// mode: 0 = round trip, 1 = compress only, 2 = decompress only (needs blobs from a prior mode-1
// call, so it compresses untimed first -- callers time mode 0 and mode 1 and subtract).
int ref_roundtrip_batch(int n_chains, const uint32_t* res_off, const uint64_t* atom_off,
                        const uint8_t* res_type, const float* xyz, const float* bfac,
                        const uint8_t* has_oxt, const float* oxt, int anchor_threshold,
                        int n_threads, int mode, uint64_t* total_bytes, uint64_t* checksum) {
    uint64_t tb = 0, cs = 0;
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : tb) reduction(^ : cs) reduction(| : err)
#endif
    for (int c = 0; c < n_chains; c++) {
        char title[32];
        int tl = snprintf(title, sizeof title, "syn_%07d", c);
        int L = (int)(res_off[c + 1] - res_off[c]);
        std::string blob;
        try {
            compress_one(res_type + res_off[c], L, xyz + 3 * atom_off[c], bfac + res_off[c],
                         has_oxt[c], oxt + 3 * c, 1, 1, 'A', title, tl, anchor_threshold, blob);
            tb += blob.size();
            if (mode != 1) {
                std::vector<AtomCoordinate> atoms;
                Foldcomp comp;
                int rc = decompress_one((const uint8_t*)blob.data(), blob.size(), 0, atoms, comp);
                if (rc != 0) err |= 1;
                uint32_t bits;
                float v = atoms.empty() ? 0.f : atoms.back().coordinate.x;
                memcpy(&bits, &v, 4);
                cs ^= bits;
            } else {
                cs ^= (uint8_t)blob[blob.size() - 1];
            }
        } catch (const std::exception&) {
            err |= 2;
        }
    }
    *total_bytes = tb;
    *checksum = cs;
    return err;
}

// PDB text of one chain given in the decoder's output layout (canonical slot order, or the -a order when
// use_alt), through the reference's own writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291).  The atoms
// are labelled the way Foldcomp::decompress labels them: residue r has residue_index idx_res + r, atom serials
// run from idx_atom, and the OXT record carries residue_index = nResidue (src/foldcomp.cpp:958-961).
// Returns the text length (the text is truncated to cap).
int64_t ref_format_pdb(const uint8_t* res_type, int L, const float* xyz, const float* bfac, int has_oxt,
                       const float* oxt, int idx_res, int idx_atom, char chain, const char* title,
                       int title_len, int use_alt, char* out, size_t cap) {
    std::vector<AtomCoordinate> atoms;
    std::string ch(1, chain);
    int ai = idx_atom;
    size_t a = 0;
    for (int r = 0; r < L; r++) {
        const TypeInfo& t = type_info(res_type[r]);
        const std::vector<std::string>* names = &t.atoms;
        auto it = aas().find(t.name3);
        if (use_alt && it != aas().end() && !it->second.altAtoms.empty()) names = &it->second.altAtoms;
        for (size_t k = 0; k < names->size(); k++, a++)
            atoms.emplace_back((*names)[k], t.name3, ch, ai++, idx_res + r, xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2], 1.0f, bfac[r]);
    }
    if (has_oxt) {
        const TypeInfo& t = type_info(res_type[L - 1]);
        atoms.emplace_back("OXT", t.name3, ch, ai++, L, oxt[0], oxt[1], oxt[2], 1.0f, bfac[L - 1]);
    }
    std::ostringstream oss;
    writeAtomCoordinatesToPDB(atoms, std::string(title, title + title_len), oss);
    const std::string s = oss.str();
    memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
    return (int64_t)s.size();
}

// The same text straight from a blob: Foldcomp::read + decompress + writeAtomCoordinatesToPDB, i.e. what
// `foldcomp decompress` writes (src/main.cpp:612-689).
int64_t ref_decompress_to_pdb(const uint8_t* fcz, size_t len, int use_alt, char* out, size_t cap) {
    std::vector<AtomCoordinate> atoms;
    Foldcomp comp;
    int rc;
    try {
        rc = decompress_one(fcz, len, use_alt, atoms, comp);
    } catch (const std::exception&) {
        return -3;
    }
    if (rc != 0) return rc;
    std::ostringstream oss;
    writeAtomCoordinatesToPDB(atoms, comp.strTitle, oss);
    const std::string s = oss.str();
    memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
    return (int64_t)s.size();
}

// Foldcomp::extract (src/foldcomp.cpp:1260-1336) on one blob: type 0 pLDDT with `digits`, type 1 sequence.
int64_t ref_extract(const uint8_t* fcz, size_t len, int type, int digits, char* out, size_t cap) {
    Foldcomp comp;
    std::istringstream iss(std::string((const char*)fcz, len));
    int rc = comp.read(iss);
    if (rc != 0) return rc;
    std::string data;
    comp.extract(data, type, digits);
    memcpy(out, data.data(), data.size() < cap ? data.size() : cap);
    return (int64_t)data.size();
}

// Foldcomp::read + Foldcomp::checkValidity (src/foldcomp.cpp:904-1036, 1492-1532) on one blob, as `foldcomp check` runs
// them (src/main.cpp:910-928): returns read()'s code; *validity = the ValidityError class.
int ref_check(const uint8_t* fcz, size_t len, int* validity) {
    Foldcomp comp;
    std::istringstream iss(std::string((const char*)fcz, len));
    *validity = 0;
    int rc = comp.read(iss);
    if (rc != 0) return rc;
    *validity = (int)comp.checkValidity();
    return 0;
}

int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
