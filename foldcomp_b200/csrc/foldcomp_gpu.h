// foldcomp_b200/csrc/foldcomp_gpu.h -- C++ host adapter over the C ABI (include/fcz_engine.h) that
// mirrors the part of class Foldcomp the reference's callers use (/root/reference/src/foldcomp.h:267-402):
//
//     Foldcomp compRes;                         FoldcompGpu compRes(engine);
//     compRes.strTitle = title;                 compRes.strTitle = title;
//     compRes.anchorThreshold = 25;             compRes.anchorThreshold = 25;
//     compRes.compress(span<AtomCoordinate>);   compRes.compress(atoms);            // src/foldcomp.cpp:562
//     compRes.writeStream(os);                  compRes.writeStream(os);            // src/foldcomp.cpp:1038
//     compRes.read(is);  (0 / -1 / -2)          compRes.read(is);                   // src/foldcomp.cpp:904
//     compRes.useAltAtomOrder = alt;            compRes.useAltAtomOrder = alt;
//     compRes.decompress(atoms);                compRes.decompress(atoms);          // src/foldcomp.cpp:779
//     compRes.write(file) / writeTar(tar, ..)   compRes.write(file) / writeTar(os, name)  // src/foldcomp.cpp:1111, 1122
//     compRes.extract(data, type, digits)       compRes.extract(data, type, digits) // src/foldcomp.cpp:1260
//     compRes.checkValidity()                   compRes.checkValidity()             // src/foldcomp.cpp:1492
//
// (integration/foldcomp_on_engine.cpp goes one step further: it implements the reference's class Foldcomp ITSELF on
// the engine, so that src/main.cpp and foldcomp/foldcomp.cxx link against it without a single changed line.)
//
// so that the CLI lambdas (src/main.cpp:438-536, 612-689) and the CPython module
// (foldcomp/foldcomp.cxx:197-220, 253-293) change only at those call sites.  The batch forms
// (compressBatch / decompressBatch) are what a batching Processor would call: one
// fcz_encode_batch / fcz_decode_batch for many chains.  Host code only; all arithmetic runs on the GPU.
#ifndef FOLDCOMP_GPU_H
#define FOLDCOMP_GPU_H

#include <iosfwd>
#include <string>
#include <vector>

#include "../../include/fcz_engine.h"

namespace fczgpu {

// same public fields as the reference's AtomCoordinate (src/atom_coordinate.h:23-55)
struct float3d {
    float x = 0, y = 0, z = 0;
};
struct AtomCoordinate {
    std::string atom, residue, chain;
    int atom_index = 0, residue_index = 0;
    float3d coordinate;
    float occupancy = 0.f, tempFactor = 0.f;
};

// One chain in the canonical slot layout of include/fcz_engine.h
struct CanonicalChain {
    std::vector<uint8_t> res_type;
    std::vector<float> bfactor, xyz;
    fcz_chain_meta meta{};
    std::string title;
};

// vector<AtomCoordinate> of ONE chain -> canonical slots (what the reference does by atom NAME at run time:
// filterBackbone src/atom_coordinate.cpp:135-143, findFirstAtomCoords src/sidechain.cpp:140-147, CA B-factor
// src/foldcomp.cpp:543-547, OXT src/foldcomp.cpp:473-481).
CanonicalChain canonicalize(const AtomCoordinate* atoms, size_t n, const std::string& title);
// decoded canonical chain -> vector<AtomCoordinate> as Foldcomp::decompress fills it (src/foldcomp.cpp:860-899)
void to_atoms(const CanonicalChain& c, bool alt_order, std::vector<AtomCoordinate>& atoms);

class Engine {  // RAII fcz_engine
public:
    explicit Engine(int device = 0);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    fcz_engine* get() const { return e_; }

private:
    fcz_engine* e_;
};

class FoldcompGpu {
public:
    explicit FoldcompGpu(Engine& eng) : eng_(eng) {}
    // knobs with the reference's names (src/foldcomp.h:289-312)
    std::string strTitle;
    int anchorThreshold = 25;
    bool useAltAtomOrder = false;
    int nResidue = 0, nAtom = 0;

    int compress(const std::vector<AtomCoordinate>& atoms);  // 0 ok, else FCZ_E_*
    int writeStream(std::ostream& os) const;
    int write(const std::string& filename) const;                          // src/foldcomp.h:377
    int writeTar(std::ostream& tar, const std::string& filename) const;     // src/foldcomp.h:382 (one archive member)
    static int writeTarEnd(std::ostream& tar);                              // the archive's closing records
    int extract(std::string& data, int type, int digits);                  // src/foldcomp.h:394: 0 pLDDT, 1 sequence
    int checkValidity();                                                    // src/foldcomp.h:401: ValidityError class, FCZ_V_*
    size_t getSize() const { return blob_.size(); }
    int read(std::istream& is);                               // 0 ok, -1 bad magic (src/foldcomp.cpp:911-915)
    int decompress(std::vector<AtomCoordinate>& atoms);       // 0 ok
    // decompress + writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291) in one engine call: the PDB text
    int decompressToPdb(std::string& text);
    const std::string& blob() const { return blob_; }

    // many chains per call
    static int compressBatch(Engine& eng, const std::vector<CanonicalChain>& chains, int anchorThreshold,
                             std::vector<std::string>& blobs, std::vector<int>& status);
    static int decompressBatch(Engine& eng, const std::vector<std::string>& blobs, bool altOrder,
                               std::vector<CanonicalChain>& chains, std::vector<int>& status);

private:
    Engine& eng_;
    std::string blob_;
};

}  // namespace fczgpu
#endif
