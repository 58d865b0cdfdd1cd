// foldcomp_b200/csrc/fcz_db.h -- host side of the batch path: a fixed-column ATOM
// parser straight to the canonical slot layout, a foldcomp-db reader/writer, and whole-database compress /
// decompress passes that feed the GPU engine with BATCHES of entries instead of one entry per OpenMP task.
//
// Reference pieces these stand in for:
//   parsePdbChain   foldcomp/foldcomp.cxx:253-293 (fixed-column ATOM records, flag 1 = no ATOM line, flag 2 =
//                   several chains) + removeAlternativePosition src/atom_coordinate.cpp:362-370 + the by-name atom
//                   lookup of the encoder (see canonicalize() in foldcomp_gpu.h)
//   DbReader        src/database_reader.cpp (index "key\toffset\tlength", lookup "key\tname\tfile", mmap'ed data)
//   DbWriter        src/database_writer.cpp:36-96 (data + .index + .lookup + .dbtype, entries sorted by key on
//                   close); entries are NUL-terminated (mmseqs convention)
//   compressDb / decompressDb   the per-entry lambdas of src/main.cpp:438-536 / 612-689 under Processor::run
//                   (src/input_processor.h:200-300), re-cut as batches over fcz_encode_batch / fcz_decode_to_pdb_*
// Host code only; no arithmetic of the codec lives here.
#ifndef FCZ_DB_H
#define FCZ_DB_H

#include <cstdint>
#include <string>
#include <vector>

#include "foldcomp_gpu.h"

namespace fczgpu {

// One single-chain PDB text -> canonical chain.  Returns 0, or the reference's flags: 1 no ATOM line, 2 more than
// one chain; 3 = an ATOM line too short to hold its columns (the reference throws std::out_of_range there).
int parsePdbChain(const char* text, size_t len, const std::string& title, CanonicalChain& out);

// One PDB text -> the units `foldcomp compress` encodes one by one (src/main.cpp:466-484): ATOM records of every chain,
// alternative positions dropped, cut into chains (identifyChains, src/atom_coordinate.cpp:469-497) and each chain into
// fragments of continuous residue numbering (identifyDiscontinousResInd, 506-530), in file order.  0, 1 (no ATOM record)
// or 3 (malformed record).
struct UnitLabel { char chain; int frag, n_frags, n_chains; };  // which chain / fragment a unit is (the CLI's output file names)
int parsePdbUnits(const char* text, size_t len, const std::string& title, std::vector<CanonicalChain>& units,
                  std::vector<UnitLabel>* labels = nullptr);

// title of the chains of one PDB text as `foldcomp compress` sets it: HEADER idCode, else TITLE records, else the file name
// (without extension); see fcz_db.cpp
std::string pdbTitle(const char* text, size_t len, const std::string& base_name);

// the reference's strtof-based field parse, with a certified fast path for plain fixed-point fields
float parseFixedFloat(const char* s, size_t n);

class DbReader {
public:
    DbReader() = default;
    ~DbReader();
    DbReader(const DbReader&) = delete;
    DbReader& operator=(const DbReader&) = delete;
    // data file `path`, index `path`.index, names from `path`.lookup when present; entries are ordered by key like the
    // reference's reader (src/database_reader.cpp:109).  Returns false on error.
    bool open(const std::string& path);
    // explicit index file; with_data = false maps no data file (index / lookup queries only)
    bool open(const std::string& path, const std::string& index_path, bool with_data);
    bool hasLookup() const { return !names_.empty(); }
    bool hasData() const { return base_ != nullptr; }
    size_t size() const { return keys_.size(); }
    uint32_t key(size_t i) const { return keys_[i]; }
    const char* data(size_t i) const { return base_ + offsets_[i]; }
    uint64_t offset(size_t i) const { return offsets_[i]; }
    uint64_t length(size_t i) const { return lengths_[i]; }  // as stored in the index (includes the trailing NUL if any)
    // payload length: the index length minus one trailing NUL byte when the entry ends with one
    uint64_t payload(size_t i) const;
    std::string name(size_t i) const;

private:
    int fd_ = -1;
    const char* base_ = nullptr;
    size_t bytes_ = 0;
    std::vector<uint32_t> keys_;
    std::vector<uint64_t> offsets_, lengths_;
    std::vector<std::string> names_;  // by entry, empty when there is no lookup file
};

class DbWriter {
public:
    DbWriter() = default;
    ~DbWriter() { close(); }
    bool open(const std::string& path);
    bool open(const std::string& path, const std::string& index_path);
    // appends data + a NUL terminator; the index records length + 1
    bool append(const char* data, size_t len, uint32_t key, const std::string& name);
    // appends data as it is (the reference's writer_append: its callers put the NUL into the data themselves)
    bool appendRaw(const char* data, size_t len, uint32_t key, const std::string& name);
    // n entries of one slab at once, entry c = base[off[c] .. off[c+1]) + a NUL terminator, written by all host threads
    // (pwrite at precomputed positions: the page-cache copy is what bounds a text database); skip[c] != 0 leaves c out
    bool appendBatch(const char* base, const uint64_t* off, size_t n, const uint32_t* keys, const std::string* names, const uint8_t* skip);
    bool close();  // writes .index / .lookup sorted by key (stable)

private:
    struct Entry { uint32_t key; uint64_t offset, length; size_t name; };
    bool put(const char* data, size_t len);  // buffered sequential write at pos_
    bool flush();
    int fd_ = -1;
    std::vector<char> buf_;
    std::string path_, index_path_;
    uint64_t pos_ = 0;       // logical end of the data file (buffered bytes included)
    std::vector<Entry> entries_;
    std::vector<std::string> names_;
};

struct DbStats {
    size_t entries = 0, failed = 0;
    uint64_t residues = 0, bytes_in = 0, bytes_out = 0;
    double seconds = 0, seconds_engine = 0;
};

// FCZ database -> PDB-text database (`foldcomp decompress --db`), entry names + ".pdb" like src/main.cpp:640-652
int decompressDb(Engine& eng, const std::string& in_db, const std::string& out_db, bool altOrder, DbStats* stats);
// PDB-text database -> FCZ database (`foldcomp compress --db`), titles = entry names without extension
int compressDb(Engine& eng, const std::string& in_db, const std::string& out_db, int anchorThreshold, DbStats* stats);

}  // namespace fczgpu

// The reference's C-style database handles (src/database_reader.h:11-27, src/database_writer.h:12-15), same names,
// signatures and return conventions, over DbReader / DbWriter: foldcomp/foldcomp.cxx:333-435 (FoldcompDatabase) and
// src/main.cpp (--db output, DatabaseProcessor input) bind to them unchanged.  `data_mode` bits as in the reference:
// 1 map the data file, 2 index cache (accepted, ignored: the index is parsed every time), 4 / 8 read the .lookup file.
// Not thread-safe, like the reference's (its callers wrap them in `omp critical`).
void* make_reader(const char* data_name, const char* index_name, int32_t data_mode);
void free_reader(void* reader);
int64_t reader_get_id(void* reader, uint32_t key);       // position in the key-sorted index, -1 when absent
const char* reader_get_data(void* reader, int64_t id);   // NULL when out of range
uint32_t reader_get_key(void* reader, int64_t id);
int64_t reader_get_length(void* reader, int64_t id);     // as in the index (includes a trailing NUL)
int64_t reader_get_offset(void* reader, int64_t id);
int64_t reader_get_size(void* reader);
uint32_t reader_lookup_entry(void* reader, const char* name);          // key of a name, UINT32_MAX when absent
const char* reader_lookup_name_alloc(void* reader, uint32_t key);      // strdup'ed name (caller frees), "" when absent
void* make_writer(const char* data_name, const char* index_name);
void free_writer(void* writer);                                        // sorts by key (stable), writes index + lookup
bool writer_append(void* writer, const char* data, size_t length, uint32_t key, const char* name);

// C entry points over the above for bindings and tests (ctypes)
extern "C" {
// parse one PDB text; arrays must hold cap_res residues / cap_atoms atoms.  Returns the parser's flag, or -1 if too small.
int fczgpu_parse_pdb(const char* text, size_t len, uint8_t* res_type, float* bfactor, float* xyz, fcz_chain_meta* meta,
                     uint32_t* n_res, uint32_t* n_atoms, uint32_t cap_res, uint32_t cap_atoms);
int fczgpu_parse_pdb_units(const char* text, size_t len, uint32_t* n_units, uint32_t* res_off, uint64_t* atom_off, uint8_t* res_type,
                           float* bfactor, float* xyz, fcz_chain_meta* meta, uint32_t cap_units, uint32_t cap_res, uint64_t cap_atoms);
float fczgpu_parse_float(const char* s, size_t n);
int fczgpu_db_copy(const char* in_db, const char* out_db);
int fczgpu_decompress_db(int device, const char* in_db, const char* out_db, int alt_order, double* stats7);
int fczgpu_compress_db(int device, const char* in_db, const char* out_db, int anchor_threshold, double* stats7);
}
#endif
