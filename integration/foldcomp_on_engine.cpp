// integration/foldcomp_on_engine.cpp -- the reference's class Foldcomp, implemented on the B200 engine.
//
// This file is the REFERENCE-SIDE BINDING of the drop-in boundary: it is compiled against the
// reference's own, unmodified headers (-I$(REF)/src; nothing of the reference is copied into this repository) and
// defines the member functions of `class Foldcomp` (/root/reference/src/foldcomp.h:267-402) that the reference's callers
// use -- the CLI lambdas (src/main.cpp:485-530 compress, 612-689 decompress, 780-860 extract, 910-928 check) and the
// CPython module (foldcomp/foldcomp.cxx:197-220 decompress, 253-293 compress, 603-671 get_data).  Linked INSTEAD of
// src/foldcomp.cpp, with every other reference source unchanged, it turns `foldcomp` (the CLI) and `foldcomp.so` (the
// Python module) into front ends of include/fcz_engine.h: the per-residue arithmetic (torsions, discretisers, bit packing,
// NeRF) runs in the CUDA kernels; src/nerf.cpp, torsion_angle.cpp, sidechain.cpp and discretizer.cpp are not on the path.
//
//   Foldcomp::compress / writeStream / write / writeTar / getSize   <- fcz_encode_batch
//   Foldcomp::read / decompress                                     <- fcz_decode_plan + fcz_decode_batch
//   Foldcomp::extract                                               <- fcz_extract_batch
//   Foldcomp::checkValidity                                         <- fcz_check_batch
//   angles after read/decompress (get_data)                         <- fcz_unpack_angles_batch
//
// The callers construct one Foldcomp per entry, from many OpenMP threads: every call leases an engine from a small
// pool (one engine serves one call at a time, include/fcz_engine.h).  That per-entry use is the compatibility path; the
// batch path for throughput is fcz_cli / compressDb / decompressDb (foldcomp_b200/csrc/fcz_db.h).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>

#include "amino_acid.h"  // must precede foldcomp.h
#include "foldcomp.h"
#include "utility.h"

#include "../include/fcz_engine.h"
#include "../foldcomp_b200/csrc/foldcomp_gpu.h"

// the reference's static table (src/foldcomp.cpp:1809); a caller may name it
const std::map<std::string, AminoAcid> Foldcomp::AAS = AminoAcid::AminoAcids();

namespace {

const char kBlobKey[] = "\x01" "fcz_blob";  // the encoded chain travels in the public strMetadata map of the object

struct Pool {
    std::mutex m;
    std::condition_variable cv;
    std::vector<fcz_engine*> idle;
    int created = 0;
    static constexpr int kMax = 8;
    fcz_engine* take() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            if (!idle.empty()) { fcz_engine* e = idle.back(); idle.pop_back(); return e; }
            if (created < kMax) {
                created++;
                lk.unlock();
                const char* dev = getenv("FCZ_DEVICE");
                fcz_engine* e = fcz_engine_create(dev ? atoi(dev) : 0, nullptr);
                if (!e) { std::cerr << "[Error] foldcomp (GPU build): no CUDA engine (fcz_engine_create failed)" << std::endl; abort(); }  // no CPU fallback
                return e;
            }
            cv.wait(lk);
        }
    }
    void give(fcz_engine* e) { { std::lock_guard<std::mutex> lk(m); idle.push_back(e); } cv.notify_one(); }
};
Pool& pool() { static Pool p; return p; }
struct Lease {
    fcz_engine* e;
    Lease() : e(pool().take()) {}
    ~Lease() { pool().give(e); }
};

std::string& blob_of(Foldcomp& f) { return f.strMetadata[kBlobKey]; }

// header + the members the callers (and writeTar / get_data) look at, from an encoded chain
int parse_into(Foldcomp& f, std::string& b) {
    if (b.size() < 4 || memcmp(b.data(), MAGICNUMBER, MAGICNUMBER_LENGTH) != 0) return -1;  // src/foldcomp.cpp:911-915
    if (b.size() < 4 + sizeof(CompressedFileHeader)) return -1;
    memcpy(&f.header, b.data() + 4, sizeof(CompressedFileHeader));
    f.nResidue = f.header.nResidue; f.nAtom = f.header.nAtom; f.idxResidue = f.header.idxResidue; f.idxAtom = f.header.idxAtom;
    f.nAllAnchor = f.header.nAnchor; f.nInnerAnchor = f.header.nAnchor - 2; f.chain = f.header.chain;
    f.nSideChainTorsion = f.header.nSideChainTorsion; f.firstResidue = f.header.firstResidue; f.lastResidue = f.header.lastResidue;
    f.lenTitle = f.header.lenTitle; f.nBackbone = 3 * f.nResidue;
    size_t o = 4 + sizeof(CompressedFileHeader);
    const size_t need = o + 4ull * f.nAllAnchor + f.lenTitle + 36ull * f.nAllAnchor + 13 + 8ull * f.nResidue + f.nSideChainTorsion + 8 + f.nResidue;
    if (b.size() < need) {
        // A blob that stops inside its last section (the B-factor bytes): `foldcomp compress --db` stores entries without
        // a terminator (src/main.cpp:510-517) while FoldcompDatabase strips one byte from every entry
        // (foldcomp/foldcomp.cxx:65,73), so the reference's own Python reader hands read() blobs that are one byte
        // short; read() (src/foldcomp.cpp:1024-1031) then decodes whatever its uninitialised buffer held there.  Here
        // the missing bytes are zeros.  Anything shorter is not a readable entry.
        if (b.size() + (size_t)f.nResidue < need) return -1;
        b.resize(need, '\0');
    }
    f.anchorIndices.resize(f.nAllAnchor);
    memcpy(f.anchorIndices.data(), b.data() + o, 4ull * f.nAllAnchor); o += 4ull * f.nAllAnchor;
    f.strTitle.assign(b.data() + o, f.lenTitle); o += f.lenTitle;
    o += 36ull * f.nAllAnchor;
    f.hasOXT = b[o];
    memcpy(&f.OXT_coords, b.data() + o + 1, 12); o += 13;
    f.residues.clear();
    for (int i = 0; i < f.nResidue; i++) f.residues.push_back(convertIntToOneLetterCode((unsigned char)b[o + 8ull * i] >> 3));
    o += 8ull * f.nResidue;
    f.sideChainAnglesDiscretized.assign((const unsigned char*)b.data() + o, (const unsigned char*)b.data() + o + f.nSideChainTorsion);
    o += f.nSideChainTorsion;
    memcpy(&f.tempFactorsDisc.min, b.data() + o, 4); memcpy(&f.tempFactorsDisc.cont_f, b.data() + o + 4, 4); o += 8;
    f.tempFactorsDiscretized.assign((const unsigned char*)b.data() + o, (const unsigned char*)b.data() + o + f.nResidue);
    return 0;
}

struct OneBlob {
    uint64_t off[2];
    fcz_blob_batch bb;
    explicit OneBlob(const std::string& b) {
        off[0] = 0; off[1] = b.size();
        memset(&bb, 0, sizeof bb);
        bb.n_chains = 1; bb.mem = FCZ_MEM_HOST; bb.blob_off = off; bb.bytes = (uint8_t*)b.data();
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------- encode side

std::vector<BackboneChain> Foldcomp::compress(const tcb::span<AtomCoordinate>& atoms) {
    std::vector<fczgpu::AtomCoordinate> in(atoms.size());
    for (size_t i = 0; i < atoms.size(); i++) {
        in[i].atom = atoms[i].atom; in[i].residue = atoms[i].residue; in[i].chain = atoms[i].chain;
        in[i].atom_index = atoms[i].atom_index; in[i].residue_index = atoms[i].residue_index;
        in[i].coordinate.x = atoms[i].coordinate.x; in[i].coordinate.y = atoms[i].coordinate.y; in[i].coordinate.z = atoms[i].coordinate.z;
        in[i].occupancy = atoms[i].occupancy; in[i].tempFactor = atoms[i].tempFactor;
    }
    const fczgpu::CanonicalChain c = fczgpu::canonicalize(in.data(), in.size(), this->strTitle);
    std::string& blob = blob_of(*this);
    blob.clear();
    this->compressedBackBone.clear();
    const uint32_t L = (uint32_t)c.res_type.size();
    if (L == 0) return this->compressedBackBone;
    uint32_t res_off[2] = {0, L}, title_off[2] = {0, (uint32_t)c.title.size()};
    uint64_t atom_off[2] = {0, c.xyz.size() / 3}, blob_off[2] = {0, 0};
    int32_t st = 0;
    fcz_chain_batch cb;
    memset(&cb, 0, sizeof cb);
    cb.n_chains = 1; cb.mem = FCZ_MEM_HOST; cb.res_off = res_off; cb.atom_off = atom_off; cb.title_off = title_off;
    cb.res_type = (uint8_t*)c.res_type.data(); cb.bfactor = (float*)c.bfactor.data(); cb.xyz = (float*)c.xyz.data();
    cb.titles = (char*)c.title.data(); cb.meta = (fcz_chain_meta*)&c.meta;
    std::string out(fcz_encode_bound(1, L, atom_off[1], c.title.size(), this->anchorThreshold), '\0');
    fcz_blob_batch ob;
    memset(&ob, 0, sizeof ob);
    ob.n_chains = 1; ob.mem = FCZ_MEM_HOST; ob.blob_off = blob_off; ob.bytes = (uint8_t*)out.data(); ob.status = &st; ob.bytes_cap = out.size();
    {
        Lease l;
        fcz_opts o;
        memset(&o, 0, sizeof o);
        o.anchor_threshold = this->anchorThreshold;
        fcz_engine_set_opts(l.e, &o);
        const int rc = fcz_encode_batch(l.e, &cb, &ob);
        if (rc != FCZ_OK || st != FCZ_OK) {
            std::cerr << "[Error] encode failed: " << (rc != FCZ_OK ? fcz_last_error(l.e) : fcz_strerror(st)) << std::endl;
            // a residue name outside the table: the reference throws std::out_of_range from AAS.at() (src/sidechain.cpp:177)
            if (st == FCZ_E_RESIDUE) throw std::out_of_range("map::at");
            return this->compressedBackBone;
        }
    }
    out.resize(blob_off[1]);
    blob.swap(out);
    parse_into(*this, blob);
    this->isPreprocessed = true; this->isCompressed = true;
    // the records as the reference's vector<BackboneChain> (convertBytesToBackboneChain, src/foldcomp.cpp:60-77)
    const size_t o_rec = 4 + sizeof(CompressedFileHeader) + 4ull * nAllAnchor + lenTitle + 36ull * nAllAnchor + 13;
    for (int i = 0; i < nResidue; i++) {
        const unsigned char* b = (const unsigned char*)blob.data() + o_rec + 8ull * i;
        BackboneChain r;
        r.residue = b[0] >> 3; r.omega = ((b[0] & 7u) << 8) | b[1]; r.psi = ((unsigned)b[2] << 4) | (b[3] >> 4); r.phi = ((b[3] & 0xFu) << 8) | b[4];
        r.ca_c_n_angle = b[5]; r.c_n_ca_angle = b[6]; r.n_ca_c_angle = b[7];
        this->compressedBackBone.push_back(r);
    }
    return this->compressedBackBone;
}

int Foldcomp::preprocess(const tcb::span<AtomCoordinate>& atoms) { this->compress(atoms); return 0; }

int Foldcomp::writeStream(std::ostream& os) {
    const std::string& b = blob_of(*this);
    os.write(b.data(), (std::streamsize)b.size());
    return 0;
}
int Foldcomp::write(std::string filename) {
    std::ofstream outfile(filename, std::ios::out | std::ios::binary);
    if (!outfile) return -1;
    return writeStream(outfile);
}
size_t Foldcomp::getSize() { return blob_of(*this).size(); }
#ifdef FOLDCOMP_EXECUTABLE
int Foldcomp::writeTar(mtar_t& tar, std::string filename, size_t size) {
    const std::string& b = blob_of(*this);
    mtar_write_file_header(&tar, filename.c_str(), (unsigned)size);
    mtar_write_data(&tar, b.data(), (unsigned)b.size());
    return 0;
}
#endif
CompressedFileHeader Foldcomp::get_header() { return this->header; }
int Foldcomp::read_header(CompressedFileHeader& h) { this->header = h; return 0; }

// ---------------------------------------------------------------------------------------------------- decode side

int Foldcomp::read(std::istream& file) {
    std::string& blob = blob_of(*this);
    blob.assign(std::istreambuf_iterator<char>(file), std::istreambuf_iterator<char>());
    return parse_into(*this, blob);
}

int Foldcomp::decompress(std::vector<AtomCoordinate>& atoms) {
    const std::string& blob = blob_of(*this);
    std::vector<fczgpu::CanonicalChain> chains;
    std::vector<int> status;
    std::vector<float> ang;
    uint64_t roff[2] = {0, 0}, total = 0;
    {
        Lease l;
        fcz_opts o;
        memset(&o, 0, sizeof o);
        o.anchor_threshold = this->anchorThreshold; o.use_alt_atom_order = this->useAltAtomOrder ? 1 : 0;
        fcz_engine_set_opts(l.e, &o);
        OneBlob ob(blob);
        uint32_t res_off[2]; uint64_t atom_off[2]; uint32_t title_off[2]; int32_t st = 0;
        fcz_chain_batch cb;
        memset(&cb, 0, sizeof cb);
        cb.n_chains = 1; cb.mem = FCZ_MEM_HOST; cb.res_off = res_off; cb.atom_off = atom_off; cb.title_off = title_off; cb.status = &st;
        fcz_sizes tot;
        if (fcz_decode_plan(l.e, &ob.bb, &cb, &tot) != FCZ_OK || st != FCZ_OK) return 1;
        fczgpu::CanonicalChain c;
        c.res_type.resize(tot.n_res); c.bfactor.resize(tot.n_res); c.xyz.resize(3 * tot.n_atoms); c.title.resize(tot.n_title_bytes);
        cb.res_type = c.res_type.data(); cb.bfactor = c.bfactor.data(); cb.xyz = c.xyz.data(); cb.titles = c.title.empty() ? nullptr : &c.title[0];
        cb.meta = &c.meta; cb.res_cap = tot.n_res; cb.atom_cap = tot.n_atoms; cb.title_cap = tot.n_title_bytes;
        if (fcz_decode_batch(l.e, &ob.bb, &cb) != FCZ_OK || st != FCZ_OK) return 1;
        std::vector<fczgpu::AtomCoordinate> out;
        fczgpu::to_atoms(c, this->useAltAtomOrder, out);
        atoms.clear();
        atoms.reserve(out.size());
        for (const auto& a : out)
            atoms.emplace_back(a.atom, a.residue, a.chain, a.atom_index, a.residue_index, a.coordinate.x, a.coordinate.y, a.coordinate.z, a.occupancy, a.tempFactor);
        this->tempFactors = c.bfactor;
        // the continuised angles Foldcomp::decompress leaves in its members (src/foldcomp.cpp:783-804): get_data reads them
        ang.resize(6 * (size_t)tot.n_res);
        if (fcz_unpack_angles_batch(l.e, &ob.bb, roff, ang.data(), tot.n_res, &total) != FCZ_OK) return 1;
    }
    const size_t L = (size_t)total;
    this->phi.resize(L); this->psi.resize(L); this->omega.resize(L);
    this->n_ca_c_angle.resize(L); this->ca_c_n_angle.resize(L); this->c_n_ca_angle.resize(L);
    this->backboneTorsionAngles.clear(); this->backboneBondAngles.clear();
    for (size_t r = 0; r < L; r++) {
        phi[r] = ang[6 * r]; psi[r] = ang[6 * r + 1]; omega[r] = ang[6 * r + 2];
        n_ca_c_angle[r] = ang[6 * r + 3]; ca_c_n_angle[r] = ang[6 * r + 4]; c_n_ca_angle[r] = ang[6 * r + 5];
    }
    for (size_t r = 0; r + 1 < L; r++) {  // src/foldcomp.cpp:788-800: psi, omega, phi / CA-C-N, C-N-CA, N-CA-C per record
        backboneTorsionAngles.push_back(psi[r]); backboneTorsionAngles.push_back(omega[r]); backboneTorsionAngles.push_back(phi[r]);
        backboneBondAngles.push_back(ca_c_n_angle[r]); backboneBondAngles.push_back(c_n_ca_angle[r]); backboneBondAngles.push_back(n_ca_c_angle[r]);
    }
    return 0;
}

int Foldcomp::continuizeTempFactors() {
    this->tempFactors.clear();
    for (unsigned int q : this->tempFactorsDiscretized) this->tempFactors.push_back((q * this->tempFactorsDisc.cont_f) + this->tempFactorsDisc.min);
    return 0;
}

int Foldcomp::extract(std::string& data, int type, int digits) {
    const std::string& blob = blob_of(*this);
    Lease l;
    OneBlob ob(blob);
    uint64_t toff[2] = {0, 0}, total = 0;
    std::string out(6ull * (size_t)std::max(this->nResidue, 1) + 16, '\0');
    fcz_text_batch tb;
    memset(&tb, 0, sizeof tb);
    tb.n_chains = 1; tb.mem = FCZ_MEM_HOST; tb.text_off = toff; tb.bytes = &out[0]; tb.bytes_cap = out.size();
    if (fcz_extract_batch(l.e, &ob.bb, type, digits, &tb, &total) != FCZ_OK) return 1;
    data.append(out.data(), (size_t)total);
    return 0;
}

ValidityError Foldcomp::checkValidity() {
    const std::string& blob = blob_of(*this);
    Lease l;
    OneBlob ob(blob);
    int32_t rs = 0, v = 0;
    if (fcz_check_batch(l.e, &ob.bb, &rs, &v) != FCZ_OK) return E_BACKBONE_COUNT_MISMATCH;
    return (ValidityError)v;
}

// host-side formatting of the callers' outputs (src/foldcomp.cpp:1223-1237)
int Foldcomp::writeFASTALike(std::ostream& os, const std::string& data) { os << ">" << this->strTitle << "\n" << data << "\n"; return 0; }
int Foldcomp::writeTSV(std::ostream& os, const std::string& data) { os << this->strTitle << "\t" << this->nResidue << "\t" << data << "\n"; return 0; }

void printValidityError(ValidityError err, std::string& filename) {  // src/foldcomp.cpp:1534-1561
    static const char* msg[] = {nullptr, "Number of backbone angles does not match header: ", "Number of sidechain angles does not match header: ",
                                "Number of temperature factors does not match header: ", "All backbone angles are empty: ",
                                "All sidechain angles are empty: ", "All temperature factors are empty: "};
    if (err == SUCCESS) return;
    if ((int)err >= 1 && (int)err <= 6) std::clog << "[Error] " << msg[(int)err] << filename << std::endl;
    else std::clog << "[Error] Unknown error: " << filename << std::endl;
}
