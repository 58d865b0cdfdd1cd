// foldcomp_b200/csrc/fcz_db.cpp -- see fcz_db.h.  Host code (g++), no codec arithmetic.
#include "fcz_db.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>
#include <future>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fcz_tables.h"

namespace fczgpu {

// ------------------------------------------------------------------------------------------ parser

// std::stof(line.substr(a, n)) of the reference = strtof on the field.  Fast path for the shape every PDB writer
// produces -- optional blanks, optional sign, digits, optional '.', digits, at most 9 digits in all: the decimal
// value m / 10^k is formed in double (m and 10^k are exact, one correctly rounded division) and rounded to float;
// that equals strtof's correctly rounded result unless the double lands exactly on the midpoint of two floats, which
// is detected from its low mantissa bits and sent to strtof.  Anything else goes to strtof as well.
float parseFixedFloat(const char* s, size_t n) {
    size_t i = 0;
    while (i < n && (s[i] == ' ' || s[i] == '\t')) i++;
    bool neg = false;
    size_t j = i;
    if (j < n && (s[j] == '-' || s[j] == '+')) { neg = s[j] == '-'; j++; }
    uint64_t m = 0;
    int nd = 0, frac = 0;
    bool dot = false, ok = true;
    for (; j < n; j++) {
        const char ch = s[j];
        if (ch >= '0' && ch <= '9') { m = m * 10 + (uint64_t)(ch - '0'); nd++; if (dot) frac++; }
        else if (ch == '.' && !dot) dot = true;
        else if (ch == ' ' || ch == '\t' || ch == '\r' || ch == 0) { break; }  // strtof stops here
        else { ok = false; break; }
    }
    if (ok && nd > 0 && nd <= 9) {
        static const double p10[10] = {1, 10, 100, 1000, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
        const double d = (double)m / p10[frac];
        uint64_t bits;
        memcpy(&bits, &d, 8);
        if ((bits & 0x1FFFFFFFull) != 0x10000000ull) {  // not a float midpoint: one more rounding is safe
            const float f = (float)d;
            return neg ? -f : f;
        }
    }
    char buf[32];
    const size_t k = n < sizeof buf - 1 ? n : sizeof buf - 1;
    memcpy(buf, s, k);
    buf[k] = 0;
    return strtof(buf, nullptr);
}

static int parse_int_field(const char* s, size_t n) {  // std::stoi on the field: blanks, sign, digits
    size_t i = 0;
    while (i < n && (s[i] == ' ' || s[i] == '\t')) i++;
    bool neg = false;
    if (i < n && (s[i] == '-' || s[i] == '+')) { neg = s[i] == '-'; i++; }
    long v = 0;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) v = v * 10 + (s[i] - '0');
    return (int)(neg ? -v : v);
}

struct RawAtom {
    uint32_t name;  // trimmed atom name, up to four characters packed little-endian (0-padded)
    uint32_t res;   // trimmed residue name, same packing
    int serial, resnum;
    float x, y, z, b;
    char chain;
};

// trim(" \t") of the reference, then the first four characters as one integer key
static uint32_t trim_key(const char* s, size_t n) {
    size_t a = 0, b = n;
    while (a < b && (s[a] == ' ' || s[a] == '\t')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) b--;
    uint32_t k = 0;
    for (size_t i = a; i < b && i < a + 4; i++) k |= (uint32_t)(uint8_t)s[i] << (8 * (i - a));
    return k;
}
static uint32_t key_of(const char* z) {
    uint32_t k = 0;
    for (int i = 0; i < 4 && z[i]; i++) k |= (uint32_t)(uint8_t)z[i] << (8 * i);
    return k;
}
struct NameKeys {
    uint32_t atom[FCZ_NUM_CODES][FCZ_MAX_ATOMS];
    uint32_t res3[FCZ_NUM_CODES];
    uint32_t ca, oxt;
    NameKeys() {
        for (int c = 0; c < FCZ_NUM_CODES; c++) {
            res3[c] = key_of(FCZ_NAME3[c]);
            for (int k = 0; k < FCZ_MAX_ATOMS; k++) atom[c][k] = key_of(FCZ_ATOM_NAME[c][k]);
        }
        ca = key_of("CA");
        oxt = key_of("OXT");
    }
};
static const NameKeys& name_keys() {
    static const NameKeys k;
    return k;
}

static int code_of_key(uint32_t r) {
    const NameKeys& nk = name_keys();
    for (int c = 0; c < FCZ_NUM_CODES; c++)
        if (nk.res3[c] == r) return FCZ_NATOMS[c] ? c : FCZ_CODE_UNK;
    return FCZ_CODE_UNK;
}

static std::string strip_ext(const std::string& n);

// ATOM records of a text in file order, alternative positions dropped (an atom named like its predecessor,
// removeAlternativePosition src/atom_coordinate.cpp:362-370).  one_chain: stop with flag 2 at a second chain id
// (foldcomp/foldcomp.cxx:266-268); otherwise every chain is read (the CLI's reader).  0, or 2 / 3 as parsePdbChain.
static int read_atom_records(const char* text, size_t len, bool one_chain, std::vector<RawAtom>& atoms) {
    atoms.clear();
    atoms.reserve(len / 81 + 1);
    char chain = 0;
    bool have_chain = false;
    size_t p = 0;
    while (p < len) {
        const char* line = text + p;
        const char* nl = (const char*)memchr(line, '\n', len - p);
        const size_t n = nl ? (size_t)(nl - line) : len - p;
        p += n + 1;
        if (n < 4 || memcmp(line, "ATOM", 4) != 0) continue;
        if (n < 22) return 3;  // substr(21, 1) would throw
        const char ch = line[21];
        if (!have_chain) { chain = ch; have_chain = true; }
        if (one_chain && ch != chain) return 2;
        if (n < 61) return 3;  // the B-factor column starts at 60
        RawAtom a;
        a.name = trim_key(line + 12, 4);
        a.res = trim_key(line + 17, 3);
        a.serial = parse_int_field(line + 6, 5);
        a.resnum = parse_int_field(line + 22, 4);
        a.x = parseFixedFloat(line + 30, 8);
        a.y = parseFixedFloat(line + 38, 8);
        a.z = parseFixedFloat(line + 46, 8);
        a.b = parseFixedFloat(line + 60, n - 60 < 6 ? n - 60 : 6);
        a.chain = ch;
        if (!atoms.empty() && atoms.back().name == a.name) continue;  // removeAlternativePosition
        atoms.push_back(a);
    }
    return 0;
}

// atoms[0..n) of ONE chain / fragment -> canonical slot layout (what Foldcomp::compress sees of them)
static void canonicalize_records(const RawAtom* atoms, size_t n, const std::string& title, CanonicalChain& out) {
    out = CanonicalChain();
    out.title = title;
    out.xyz.reserve(3 * n + 64);
    out.res_type.reserve(n / 4 + 8);
    out.bfactor.reserve(n / 4 + 8);
    out.meta.n_atom = (uint16_t)n;
    out.meta.idx_residue = (uint16_t)atoms[0].resnum;
    out.meta.idx_atom = (uint16_t)atoms[0].serial;
    out.meta.chain = (uint8_t)atoms[0].chain;
    const NameKeys& nk = name_keys();
    if (atoms[n - 1].name == nk.oxt) {  // src/foldcomp.cpp:473-481
        out.meta.has_oxt = 1;
        out.meta.oxt[0] = atoms[n - 1].x; out.meta.oxt[1] = atoms[n - 1].y; out.meta.oxt[2] = atoms[n - 1].z;
    }
    size_t i = 0;
    while (i < n) {  // splitAtomByResidue (src/atom_coordinate.cpp:304-328): the last atom joins the current residue
        size_t j = i + 1;
        while (j < n && (atoms[j].resnum == atoms[j - 1].resnum || j == n - 1)) j++;
        const int code = code_of_key(atoms[i].res);
        out.res_type.push_back((uint8_t)code);
        for (int k = 0; k < FCZ_NATOMS[code]; k++) {
            const uint32_t want = nk.atom[code][k];
            float x = 0, y = 0, z = 0;  // findFirstAtomCoords: a missing atom reads as (0,0,0)
            // the FIRST atom of that name in the residue; files written in table order hit it at position k, but an
            // earlier duplicate name must still win, so the prefix is checked before the shortcut is taken
            size_t hit = j;
            if (i + k < j && atoms[i + k].name == want) {
                hit = i + k;
                for (size_t a = i; a < i + k; a++)
                    if (atoms[a].name == want) { hit = a; break; }
            } else {
                for (size_t a = i; a < j; a++)
                    if (atoms[a].name == want) { hit = a; break; }
            }
            if (hit < j) { x = atoms[hit].x; y = atoms[hit].y; z = atoms[hit].z; }
            out.xyz.push_back(x); out.xyz.push_back(y); out.xyz.push_back(z);
        }
        float bf = 0.f;
        for (size_t a = i; a < j; a++)
            if (atoms[a].name == nk.ca) { bf = atoms[a].b; break; }
        out.bfactor.push_back(bf);
        i = j;
    }
}

int parsePdbChain(const char* text, size_t len, const std::string& title, CanonicalChain& out) {
    out = CanonicalChain();
    out.title = title;
    std::vector<RawAtom> atoms;
    const int flag = read_atom_records(text, len, true, atoms);
    if (flag) return flag;
    if (atoms.empty()) return 1;
    canonicalize_records(atoms.data(), atoms.size(), title, out);
    return 0;
}

// The title `foldcomp compress` gives the chains of one PDB text (src/main.cpp:466-467 over StructureReader::updateStructure,
// src/structure_reader.cpp:31-46, over gemmi's record reader, lib/gemmi/pdb.hpp:483-498): the HEADER record's idCode
// (columns 63-66) when the line is long enough and the code is not blank; else the text of the TITLE records (from column
// 11 on, right-trimmed, continuation lines appended as they stand); else the file name -- and a title equal to the file name
// becomes the name without its extension.  Records after the first ATOM / HETATM line are not looked at (gemmi would).
std::string pdbTitle(const char* text, size_t len, const std::string& base_name) {
    std::string entry_id, title;
    size_t p = 0;
    auto rtrim = [](std::string v) {
        const size_t last = v.find_last_not_of(" \r\n\t");
        return v.substr(0, last == std::string::npos ? 0 : last + 1);
    };
    while (p < len) {
        const char* line = text + p;
        const char* nl = (const char*)memchr(line, '\n', len - p);
        const size_t n = nl ? (size_t)(nl - line) : len - p;  // without the newline
        const size_t glen = n + (nl ? 1 : 0);                  // gemmi's `len` counts the newline it copied
        p += n + 1;
        if (n >= 4 && (memcmp(line, "ATOM", 4) == 0 || memcmp(line, "HETA", 4) == 0)) break;
        if (n >= 6 && memcmp(line, "HEADER", 6) == 0) {
            if (glen > 66) { const std::string id = rtrim(std::string(line + 62, 4)); if (!id.empty()) entry_id = id; }
        } else if (n >= 5 && memcmp(line, "TITLE", 5) == 0) {
            if (glen > 10) title += rtrim(std::string(line + 10, glen - 10 - 1));
        }
    }
    std::string t = !entry_id.empty() ? entry_id : (!title.empty() ? title : base_name);
    if (t == base_name) t = strip_ext(base_name);
    return t;
}

int parsePdbUnits(const char* text, size_t len, const std::string& title, std::vector<CanonicalChain>& units, std::vector<UnitLabel>* labels) {
    units.clear();
    if (labels) labels->clear();
    std::vector<RawAtom> atoms;
    const int flag = read_atom_records(text, len, false, atoms);
    if (flag) return flag;
    if (atoms.empty()) return 1;
    const NameKeys& nk = name_keys();
    const uint32_t kN = key_of("N");
    (void)nk;
    const size_t n = atoms.size();
    // identifyChains (src/atom_coordinate.cpp:469-497): a new chain starts where the chain id changes -- at that atom when
    // it is an N, else at the next N atom (the atoms in between belong to nobody).  (Without any further N atom the
    // reference never leaves its loop; here the rest of the text is dropped.)
    std::vector<std::pair<size_t, size_t>> chains;
    size_t start = 0;
    for (size_t i = 1; i < n; i++) {
        if (atoms[i].chain == atoms[i - 1].chain) continue;
        size_t j = i;
        while (j < n && atoms[j].name != kN) j++;
        chains.emplace_back(start, i);
        start = j;
        if (j >= n) break;
        i = j;
    }
    if (start < n) chains.emplace_back(start, n);
    for (const auto& ch : chains) {
        // identifyDiscontinousResInd (src/atom_coordinate.cpp:506-530): fragments start at N atoms; a fragment ends where
        // the next N atom's residue number exceeds the previous N atom's by more than one.  The first fragment starts at
        // the chain's first N atom.  (A chain without N atoms is undefined behaviour there; here it stays one fragment.)
        size_t fstart = ch.first;
        bool have = false;
        int prev = 0;
        std::vector<std::pair<size_t, size_t>> frags;
        for (size_t i = ch.first; i < ch.second; i++) {
            if (atoms[i].name != kN) continue;
            if (!have) { fstart = i; have = true; }
            else if (atoms[i].resnum - prev > 1) { frags.emplace_back(fstart, i); fstart = i; }
            prev = atoms[i].resnum;
        }
        frags.emplace_back(fstart, ch.second);
        int j = 0;
        for (const auto& f : frags) {
            if (f.second > f.first) {
                units.emplace_back();
                canonicalize_records(atoms.data() + f.first, f.second - f.first, title, units.back());
                if (labels) labels->push_back({atoms[ch.first].chain, j, (int)frags.size(), (int)chains.size()});
            }
            j++;
        }
    }
    return units.empty() ? 1 : 0;
}

// ------------------------------------------------------------------------------------------ reader

DbReader::~DbReader() {
    if (base_) munmap((void*)base_, bytes_);
    if (fd_ >= 0) ::close(fd_);
}

static bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    out.clear();
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    fclose(f);
    return true;
}

bool DbReader::open(const std::string& path) { return open(path, path + ".index", true); }

bool DbReader::open(const std::string& path, const std::string& index_path, bool with_data) {
    std::string idx;
    if (!read_file(index_path, idx)) return false;
    const char* p = idx.c_str();
    while (*p) {
        char* e;
        const unsigned long k = strtoul(p, &e, 10);
        if (e == p) break;
        const unsigned long long off = strtoull(e, &e, 10), ln = strtoull(e, &e, 10);
        keys_.push_back((uint32_t)k); offsets_.push_back(off); lengths_.push_back(ln);
        p = e;
        while (*p == '\n' || *p == '\r' || *p == ' ' || *p == '\t') p++;
    }
    {   // entries by key, like the reference's reader (std::sort by id, src/database_reader.cpp:109); stable for equal keys
        std::vector<size_t> perm(keys_.size());
        for (size_t i = 0; i < perm.size(); i++) perm[i] = i;
        std::stable_sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return keys_[x] < keys_[y]; });
        std::vector<uint32_t> k2(keys_.size());
        std::vector<uint64_t> o2(keys_.size()), l2(keys_.size());
        for (size_t i = 0; i < perm.size(); i++) { k2[i] = keys_[perm[i]]; o2[i] = offsets_[perm[i]]; l2[i] = lengths_[perm[i]]; }
        keys_.swap(k2); offsets_.swap(o2); lengths_.swap(l2);
    }
    if (with_data) {  // (without: index / lookup queries only, the reference's reader without DB_READER_USE_DATA)
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) return false;
        struct stat st;
        if (fstat(fd_, &st) != 0) return false;
        bytes_ = (size_t)st.st_size;
        if (bytes_) {
            void* m = mmap(nullptr, bytes_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) return false;
            base_ = (const char*)m;
        }
        for (size_t i = 0; i < keys_.size(); i++)  // no sum: offset + length of a malformed row may wrap
            if (lengths_[i] > bytes_ || offsets_[i] > bytes_ - lengths_[i]) return false;
    }
    std::string lk;
    if (read_file(path + ".lookup", lk)) {
        names_.assign(keys_.size(), std::string());
        std::vector<std::pair<uint32_t, std::string>> rows;
        size_t a = 0;
        while (a < lk.size()) {
            size_t b = lk.find('\n', a);
            if (b == std::string::npos) b = lk.size();
            const size_t t1 = lk.find('\t', a);
            if (t1 != std::string::npos && t1 < b) {
                size_t t2 = lk.find('\t', t1 + 1);
                if (t2 == std::string::npos || t2 > b) t2 = b;
                rows.emplace_back((uint32_t)strtoul(lk.c_str() + a, nullptr, 10), lk.substr(t1 + 1, t2 - t1 - 1));
            }
            a = b + 1;
        }
        std::sort(rows.begin(), rows.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
        for (size_t i = 0; i < keys_.size(); i++) {
            auto it = std::lower_bound(rows.begin(), rows.end(), keys_[i], [](const auto& x, uint32_t k) { return x.first < k; });
            if (it != rows.end() && it->first == keys_[i]) names_[i] = it->second;
        }
    }
    return true;
}

uint64_t DbReader::payload(size_t i) const {
    const uint64_t n = lengths_[i];
    return (n && base_[offsets_[i] + n - 1] == 0) ? n - 1 : n;
}

std::string DbReader::name(size_t i) const {
    if (i < names_.size() && !names_[i].empty()) return names_[i];
    return std::to_string(keys_[i]);
}

// ------------------------------------------------------------------------------------------ writer

bool DbWriter::open(const std::string& path) { return open(path, path + ".index"); }

bool DbWriter::open(const std::string& path, const std::string& index_path) {
    path_ = path;
    index_path_ = index_path;
    fd_ = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd_ < 0) return false;
    FILE* t = fopen((path + ".dbtype").c_str(), "wb");
    if (!t) { ::close(fd_); fd_ = -1; return false; }
    const int type = 12;  // generic dbtype, src/database_writer.cpp:51-55
    const bool ok = fwrite(&type, sizeof type, 1, t) == 1;
    if (fclose(t) != 0 || !ok) { ::close(fd_); fd_ = -1; return false; }
    pos_ = 0;
    buf_.clear();
    buf_.reserve(4u << 20);
    return true;
}

static bool pwrite_all(int fd, const char* p, size_t n, uint64_t at) {
    while (n) {
        const ssize_t w = ::pwrite(fd, p, n, (off_t)at);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w; at += (uint64_t)w;
    }
    return true;
}

bool DbWriter::flush() {
    if (buf_.empty()) return true;
    const bool ok = pwrite_all(fd_, buf_.data(), buf_.size(), pos_ - buf_.size());
    buf_.clear();
    return ok;
}

bool DbWriter::put(const char* data, size_t len) {
    if (len >= (1u << 20)) {  // large payloads go straight to the file
        if (!flush() || !pwrite_all(fd_, data, len, pos_)) return false;
        pos_ += len;
        return true;
    }
    if (buf_.size() + len > (4u << 20) && !flush()) return false;
    buf_.insert(buf_.end(), data, data + len);
    pos_ += len;
    return true;
}

bool DbWriter::append(const char* data, size_t len, uint32_t key, const std::string& name) {
    if (fd_ < 0) return false;
    const uint64_t at = pos_;
    const char nul = 0;
    if (!put(data, len) || !put(&nul, 1)) return false;
    names_.push_back(name);
    entries_.push_back({key, at, (uint64_t)len + 1, names_.size() - 1});
    return true;
}

bool DbWriter::appendRaw(const char* data, size_t len, uint32_t key, const std::string& name) {
    if (fd_ < 0) return false;
    const uint64_t at = pos_;
    if (!put(data, len)) return false;
    names_.push_back(name);
    entries_.push_back({key, at, (uint64_t)len, names_.size() - 1});
    return true;
}

bool DbWriter::appendBatch(const char* base, const uint64_t* off, size_t n, const uint32_t* keys, const std::string* names, const uint8_t* skip) {
    if (fd_ < 0 || !flush()) return false;
    std::vector<uint64_t> at(n + 1);
    at[0] = pos_;
    for (size_t c = 0; c < n; c++) at[c + 1] = at[c] + ((skip && skip[c]) ? 0 : off[c + 1] - off[c] + 1);
    int bad = 0;
    // Buffered write()s to ONE file are serialised by the file system (the inode lock), so pwrite from many threads runs
    // at little more than the speed of one; stores through a shared mapping fault their pages in concurrently but pay a
    // fault per page -- measured on the B200 box's ext4 volume (profiles/r02_v14b_cli_e2e_50k.json): 11 GB of text in
    // 3.8 s by pwrite, 4.2 s through the mapping.  pwrite stays the default; FCZ_DB_WRITE=mmap selects the mapping.
    const char* wmode = getenv("FCZ_DB_WRITE");
    const bool use_mmap = wmode && strcmp(wmode, "mmap") == 0;
    bool mapped = false;
    if (use_mmap && at[n] > pos_) {
        const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE), m0 = pos_ & ~(page - 1);
        if (ftruncate(fd_, (off_t)at[n]) == 0) {
            void* p = mmap(nullptr, at[n] - m0, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, (off_t)m0);
            if (p != MAP_FAILED) {
                char* dst = (char*)p - m0;  // dst + file offset
#pragma omp parallel for schedule(dynamic, 16)
                for (size_t c = 0; c < n; c++) {
                    if (skip && skip[c]) continue;
                    const size_t len = (size_t)(off[c + 1] - off[c]);
                    memcpy(dst + at[c], base + off[c], len);
                    dst[at[c] + len] = 0;
                }
                mapped = munmap(p, at[n] - m0) == 0;
                if (!mapped) return false;
            }
        }
    }
    if (!mapped) {
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
        for (size_t c = 0; c < n; c++) {
            if (skip && skip[c]) continue;
            const size_t len = (size_t)(off[c + 1] - off[c]);
            const char nul = 0;
            if (!pwrite_all(fd_, base + off[c], len, at[c]) || !pwrite_all(fd_, &nul, 1, at[c] + len)) bad |= 1;
        }
    }
    if (bad) return false;
    for (size_t c = 0; c < n; c++) {
        if (skip && skip[c]) continue;
        names_.push_back(names[c]);
        entries_.push_back({keys[c], at[c], at[c + 1] - at[c], names_.size() - 1});
    }
    pos_ = at[n];
    return true;
}

bool DbWriter::close() {
    if (fd_ < 0) return true;
    bool ok = flush();
    ok &= ::close(fd_) == 0;  // a full disk shows up here at the latest: never report a truncated database as written
    fd_ = -1;
    std::stable_sort(entries_.begin(), entries_.end(), [](const Entry& a, const Entry& b) { return a.key < b.key; });
    FILE* idx = fopen(index_path_.c_str(), "w");
    FILE* lk = fopen((path_ + ".lookup").c_str(), "w");
    if (!idx || !lk) {
        if (idx) fclose(idx);
        if (lk) fclose(lk);
        return false;
    }
    for (const Entry& e : entries_) {
        ok &= fprintf(idx, "%u\t%llu\t%llu\n", e.key, (unsigned long long)e.offset, (unsigned long long)e.length) > 0;
        ok &= fprintf(lk, "%u\t%s\t0\n", e.key, names_[e.name].c_str()) > 0;
    }
    ok &= fclose(idx) == 0;
    ok &= fclose(lk) == 0;
    entries_.clear();
    return ok;
}

// ------------------------------------------------------------------------------------- whole-db passes

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static std::string strip_ext(const std::string& n) {
    const size_t d = n.find_last_of('.');
    return d == std::string::npos ? n : n.substr(0, d);
}

int decompressDb(Engine& eng, const std::string& in_db, const std::string& out_db, bool altOrder, DbStats* stats) {
    // Per batch: blobs up, decode + text on the GPU, text back into one of two pinned buffers; while the GPU works on the
    // next batch a writer thread puts the previous one into the data file with all host cores (DbWriter::appendBatch).
    const double t0 = now_s();
    DbReader rd;
    if (!rd.open(in_db)) return FCZ_E_ARG;
    DbWriter wr;
    if (!wr.open(out_db)) return FCZ_E_ARG;
    DbStats s;
    fcz_opts o{25, altOrder ? 1 : 0, nullptr, 0};
    int rc = fcz_engine_set_opts(eng.get(), &o);
    if (rc) return rc;
    const size_t n_all = rd.size();
    const uint64_t kBatchBytes = 8ull << 20;  // ~8 MB of FCZ per engine call (~330 MB of text)
    struct Slot {
        char* text = nullptr; uint64_t cap = 0;
        std::vector<uint64_t> text_off;
        std::vector<uint32_t> keys;
        std::vector<std::string> names;
        std::vector<uint8_t> skip;
        std::thread writer;
        bool ok = true;
    } slot[2];
    struct Free { Slot* s; ~Free() { for (int i = 0; i < 2; i++) { if (s[i].writer.joinable()) s[i].writer.join(); if (s[i].text) fcz_host_free(s[i].text); } } } guard{slot};
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> blob_off;
    std::vector<int32_t> status;
    size_t i0 = 0;
    for (int k = 0; i0 < n_all; k ^= 1) {
        size_t i1 = i0;
        uint64_t sum = 0;
        while (i1 < n_all && (i1 == i0 || sum + rd.length(i1) <= kBatchBytes)) sum += rd.length(i1++);
        const uint32_t n = (uint32_t)(i1 - i0);
        Slot& sl = slot[k];
        if (sl.writer.joinable()) sl.writer.join();  // the batch that used this buffer two rounds ago
        if (!sl.ok) return FCZ_E_ARG;
        blob_off.assign(n + 1, 0);
        for (uint32_t c = 0; c < n; c++) blob_off[c + 1] = blob_off[c] + rd.length(i0 + c);
        // entries of one database usually lie back to back: hand the mapped file to the engine as it is
        bool contiguous = true;
        for (uint32_t c = 0; c + 1 < n; c++) contiguous &= rd.offset(i0 + c) + rd.length(i0 + c) == rd.offset(i0 + c + 1);
        const uint8_t* src = (const uint8_t*)rd.data(i0);
        if (!contiguous) {
            bytes.resize(blob_off[n] + 1);
            for (uint32_t c = 0; c < n; c++) memcpy(bytes.data() + blob_off[c], rd.data(i0 + c), rd.length(i0 + c));
            src = bytes.data();
        }
        fcz_blob_batch in{};
        in.n_chains = n; in.mem = FCZ_MEM_HOST; in.blob_off = blob_off.data(); in.bytes = const_cast<uint8_t*>(src);
        sl.text_off.assign(n + 1, 0);
        status.assign(n + 1, 0);
        fcz_text_batch out{};
        out.n_chains = n; out.mem = FCZ_MEM_HOST; out.text_off = sl.text_off.data(); out.status = status.data();
        uint64_t total = 0;
        const double g0 = now_s();
        if ((rc = fcz_decode_to_pdb_plan(eng.get(), &in, &out, &total))) return rc;
        if (sl.cap < total + 1) {
            if (sl.text) fcz_host_free(sl.text);
            sl.cap = total + total / 8 + 4096;
            sl.text = (char*)fcz_host_alloc(sl.cap);
            if (!sl.text) return FCZ_E_CUDA;
        }
        out.bytes = sl.text; out.bytes_cap = sl.cap;
        if ((rc = fcz_decode_to_pdb_batch(eng.get(), &in, &out))) return rc;
        s.seconds_engine += now_s() - g0;
        sl.keys.resize(n); sl.names.resize(n); sl.skip.assign(n, 0);
        for (uint32_t c = 0; c < n; c++) {
            s.entries++;
            s.bytes_in += rd.length(i0 + c);
            sl.keys[c] = rd.key(i0 + c);
            sl.names[c] = strip_ext(rd.name(i0 + c)) + ".pdb";
            if (status[c] != FCZ_OK) { s.failed++; sl.skip[c] = 1; continue; }
            const uint8_t* b = src + blob_off[c];
            s.residues += (uint64_t)b[4] | (uint64_t)b[5] << 8;  // CompressedFileHeader.nResidue
            s.bytes_out += sl.text_off[c + 1] - sl.text_off[c];
        }
        // batches are appended in order: the previous writer (other slot) must be done before this one starts
        Slot& prev = slot[k ^ 1];
        if (prev.writer.joinable()) prev.writer.join();
        if (!prev.ok) return FCZ_E_ARG;
        sl.writer = std::thread([&wr, &sl, n]() { sl.ok = wr.appendBatch(sl.text, sl.text_off.data(), n, sl.keys.data(), sl.names.data(), sl.skip.data()); });
        i0 = i1;
    }
    for (int k = 0; k < 2; k++) {
        if (slot[k].writer.joinable()) slot[k].writer.join();
        if (!slot[k].ok) return FCZ_E_ARG;
    }
    if (!wr.close()) return FCZ_E_ARG;
    s.seconds = now_s() - t0;
    if (stats) *stats = s;
    return FCZ_OK;
}

int compressDb(Engine& eng, const std::string& in_db, const std::string& out_db, int anchorThreshold, DbStats* stats) {
    // PDB text goes to the GPU as it is: the ATOM parser (fcz_parse.h, k_parse_*) and the encoder run there, only the FCZ
    // blobs (about 1/40 of the text) come back.  The host's share is one parallel copy of the mapped text into a pinned
    // buffer -- without each entry's NUL terminator; the copy of batch k+1 runs while the GPU works on batch k -- and
    // the database writer.  FCZ_HOST_PARSER=1 restores the per-entry host parser (parsePdbChain under OpenMP) for A/B runs.
    const double t0 = now_s();
    DbReader rd;
    if (!rd.open(in_db)) return FCZ_E_ARG;
    DbWriter wr;
    if (!wr.open(out_db)) return FCZ_E_ARG;
    DbStats s;
    const size_t n_all = rd.size();
    const char* hp = getenv("FCZ_HOST_PARSER");
    const bool host_parser = hp && atoi(hp) != 0;
    // Outputs like the reference CLI's (src/main.cpp:438-536): an entry whose ATOM records hold several chains or breaks in
    // the residue numbering yields one FCZ entry per chain / fragment; every output carries the entry's base name (no
    // extension, 448-449) and the next key of a running counter (514-517; there in completion order, here in input order).
    uint32_t next_key = 0;
    fcz_opts o{anchorThreshold, 0, nullptr, 0};
    int rc = fcz_engine_set_opts(eng.get(), &o);
    if (rc) return rc;
    if (host_parser) {
        const uint64_t kBatchBytes = 512ull << 20;  // PDB text per engine call (~80 MB of coordinates)
        size_t i0 = 0;
        while (i0 < n_all) {
            size_t i1 = i0;
            uint64_t sum = 0;
            while (i1 < n_all && (i1 == i0 || sum + rd.length(i1) <= kBatchBytes)) sum += rd.length(i1++);
            const size_t n = i1 - i0;
            std::vector<std::vector<CanonicalChain>> units(n);
            std::vector<int> flag(n, 0);
#pragma omp parallel for schedule(dynamic, 8)
            for (size_t c = 0; c < n; c++)
                flag[c] = parsePdbUnits(rd.data(i0 + c), rd.payload(i0 + c), pdbTitle(rd.data(i0 + c), rd.payload(i0 + c), rd.name(i0 + c)), units[c]);
            std::vector<CanonicalChain> good;
            std::vector<size_t> which;
            for (size_t c = 0; c < n; c++) {
                s.entries++;
                s.bytes_in += rd.length(i0 + c);
                if (flag[c] != 0) { s.failed++; continue; }
                for (auto& u : units[c]) { good.push_back(std::move(u)); which.push_back(c); }
            }
            std::vector<std::string> blobs;
            std::vector<int> st;
            const double g0 = now_s();
            rc = FoldcompGpu::compressBatch(eng, good, anchorThreshold, blobs, st);
            s.seconds_engine += now_s() - g0;
            if (rc) return rc;
            for (size_t g = 0; g < good.size(); g++) {
                if (st[g] != FCZ_OK) { s.failed++; continue; }
                s.residues += good[g].res_type.size();
                s.bytes_out += blobs[g].size();
                if (!wr.appendRaw(blobs[g].data(), blobs[g].size(), next_key++, strip_ext(rd.name(i0 + which[g])))) return FCZ_E_ARG;
            }
            i0 = i1;
        }
        if (!wr.close()) return FCZ_E_ARG;
        s.seconds = now_s() - t0;
        if (stats) *stats = s;
        return FCZ_OK;
    }
    // ---- GPU parser, two pinned text buffers
    const uint64_t kBatchBytes = 256ull << 20;  // PDB text per engine call
    struct Slot {
        char* text = nullptr; uint64_t cap = 0;
        size_t i0 = 0, i1 = 0;
        std::vector<uint64_t> text_off;
        std::thread copier;
    } slot[2];
    uint8_t* pin_blob = nullptr;
    uint64_t pin_blob_cap = 0;
    struct Free { Slot* s; uint8_t*& b; ~Free() { for (int i = 0; i < 2; i++) { if (s[i].copier.joinable()) s[i].copier.join(); if (s[i].text) fcz_host_free(s[i].text); } if (b) fcz_host_free(b); } } guard{slot, pin_blob};
    // stage(k, i0): picks the entries of the batch that starts at i0, sizes the buffer, starts the copy; returns false on error
    auto stage = [&](Slot& sl, size_t i0) -> bool {
        size_t i1 = i0;
        uint64_t sum = 0;
        while (i1 < n_all && (i1 == i0 || sum + rd.length(i1) <= kBatchBytes)) sum += rd.length(i1++);
        sl.i0 = i0; sl.i1 = i1;
        const size_t n = i1 - i0;
        sl.text_off.assign(n + 1, 0);
        for (size_t c = 0; c < n; c++) sl.text_off[c + 1] = sl.text_off[c] + rd.payload(i0 + c);
        if (sl.text_off[n] + 16 > sl.cap) {
            if (sl.text) fcz_host_free(sl.text);
            sl.cap = std::max<uint64_t>(sl.text_off[n] + sl.text_off[n] / 16 + 4096, kBatchBytes + (kBatchBytes >> 4));
            sl.text = (char*)fcz_host_alloc(sl.cap);
            if (!sl.text) return false;
        }
        sl.copier = std::thread([&rd, &sl, n]() {
#pragma omp parallel for schedule(static)
            for (size_t c = 0; c < n; c++) memcpy(sl.text + sl.text_off[c], rd.data(sl.i0 + c), sl.text_off[c + 1] - sl.text_off[c]);
        });
        return true;
    };
    if (n_all && !stage(slot[0], 0)) return FCZ_E_CUDA;
    for (int k = 0; n_all && slot[k].i0 < n_all; k ^= 1) {
        Slot& sl = slot[k];
        sl.copier.join();
        const size_t i0 = sl.i0, i1 = sl.i1, n = i1 - i0;
        if (i1 < n_all) { if (!stage(slot[k ^ 1], i1)) return FCZ_E_CUDA; }
        else slot[k ^ 1].i0 = n_all;
        std::vector<uint64_t> blob_off(n + 1, 0);
        std::vector<uint32_t> title_off(n + 1, 0);
        std::vector<std::string> names(n);
        std::string titles;
        std::vector<std::string> title_of(n);
#pragma omp parallel for schedule(static)
        for (size_t c = 0; c < n; c++) title_of[c] = pdbTitle(rd.data(i0 + c), rd.payload(i0 + c), rd.name(i0 + c));
        for (size_t c = 0; c < n; c++) {
            names[c] = strip_ext(rd.name(i0 + c));
            titles += title_of[c];
            title_off[c + 1] = (uint32_t)titles.size();
        }
        const uint64_t n_text = sl.text_off[n];
        const uint64_t blob_cap = n_text / 16 + 512 * n + titles.size() + 4096;  // FCZ is ~1/40 of its text; retried when short
        if (blob_cap > pin_blob_cap) {
            if (pin_blob) fcz_host_free(pin_blob);
            pin_blob_cap = blob_cap + blob_cap / 8;
            pin_blob = (uint8_t*)fcz_host_alloc(pin_blob_cap);
            if (!pin_blob) return FCZ_E_CUDA;
        }
        std::vector<int32_t> status(n + 1, 0);
        fcz_text_batch in{};
        in.n_chains = (uint32_t)n; in.mem = FCZ_MEM_HOST; in.text_off = sl.text_off.data(); in.bytes = sl.text; in.bytes_cap = n_text;
        fcz_blob_batch out{};
        out.n_chains = (uint32_t)n; out.mem = FCZ_MEM_HOST; out.blob_off = blob_off.data(); out.bytes = pin_blob; out.bytes_cap = pin_blob_cap;
        out.status = status.data();
        uint64_t total = 0;
        const double g0 = now_s();
        rc = fcz_encode_pdb_text_batch(eng.get(), &in, title_off.data(), titles.data(), &out, &total);
        if (rc == FCZ_E_CAPACITY) {
            fcz_host_free(pin_blob);
            pin_blob_cap = total + 4096;
            pin_blob = (uint8_t*)fcz_host_alloc(pin_blob_cap);
            if (!pin_blob) return FCZ_E_CUDA;
            out.bytes = pin_blob; out.bytes_cap = pin_blob_cap;
            rc = fcz_encode_pdb_text_batch(eng.get(), &in, title_off.data(), titles.data(), &out, &total);
        }
        s.seconds_engine += now_s() - g0;
        if (rc) return rc;
        // What the GPU parser hands back to the host: entries with several chains or fragments (the units are cut here,
        // parsePdbUnits) and entries with a numeric field outside the GPU grammar (exponents, hex floats: no PDB writer
        // emits them; the host parser goes through strtof).  Their blobs are made by one more engine call.
        std::vector<CanonicalChain> odd;
        std::vector<size_t> which;
        std::vector<std::string> blobs_host;
        std::vector<int> st_host;
        for (size_t c = 0; c < n; c++) {
            if (status[c] != FCZ_E_PARSE_NUMBER && status[c] != FCZ_E_PARSE_CHAINS && status[c] != FCZ_E_PARSE_GAPS) continue;
            std::vector<CanonicalChain> units;
            if (parsePdbUnits(rd.data(i0 + c), rd.payload(i0 + c), title_of[c], units) != 0) continue;
            for (auto& u : units) { odd.push_back(std::move(u)); which.push_back(c); }
        }
        if (!odd.empty() && (rc = FoldcompGpu::compressBatch(eng, odd, anchorThreshold, blobs_host, st_host))) return rc;
        size_t w = 0;
        for (size_t c = 0; c < n; c++) {
            s.entries++;
            s.bytes_in += rd.length(i0 + c);
            if (w < which.size() && which[w] == c) {
                bool any = false;
                for (; w < which.size() && which[w] == c; w++) {
                    if (st_host[w] != FCZ_OK) continue;
                    any = true;
                    s.residues += odd[w].res_type.size();
                    s.bytes_out += blobs_host[w].size();
                    if (!wr.appendRaw(blobs_host[w].data(), blobs_host[w].size(), next_key++, names[c])) return FCZ_E_ARG;
                }
                if (!any) s.failed++;
                continue;
            }
            const char* data = (const char*)pin_blob + blob_off[c];
            const size_t len = (size_t)(blob_off[c + 1] - blob_off[c]);
            if (status[c] != FCZ_OK || len < 6) { s.failed++; continue; }
            s.residues += (uint64_t)(uint8_t)data[4] | (uint64_t)(uint8_t)data[5] << 8;  // CompressedFileHeader.nResidue
            s.bytes_out += len;
            if (!wr.appendRaw(data, len, next_key++, names[c])) return FCZ_E_ARG;  // no terminator, like `foldcomp compress --db` (src/main.cpp:510-517)
        }
    }
    if (!wr.close()) return FCZ_E_ARG;
    s.seconds = now_s() - t0;
    if (stats) *stats = s;
    return FCZ_OK;
}

}  // namespace fczgpu

using namespace fczgpu;

extern "C" int fczgpu_parse_pdb(const char* text, size_t len, uint8_t* res_type, float* bfactor, float* xyz, fcz_chain_meta* meta,
                                uint32_t* n_res, uint32_t* n_atoms, uint32_t cap_res, uint32_t cap_atoms) {
    CanonicalChain c;
    const int flag = parsePdbChain(text, len, "", c);
    if (flag) return flag;
    *n_res = (uint32_t)c.res_type.size();
    *n_atoms = (uint32_t)(c.xyz.size() / 3);
    if (*n_res > cap_res || *n_atoms > cap_atoms) return -1;
    memcpy(res_type, c.res_type.data(), c.res_type.size());
    memcpy(bfactor, c.bfactor.data(), 4 * c.bfactor.size());
    memcpy(xyz, c.xyz.data(), 4 * c.xyz.size());
    *meta = c.meta;
    return 0;
}

// test hook for DbWriter::appendBatch: a database of n entries cut from one slab (entry c = base[off[c] .. off[c+1]), keys
// 100 + c, names "e<c>"), preceded and followed by one ordinary append; skip may be NULL
extern "C" int fczgpu_db_write_batch(const char* path, const char* base, const uint64_t* off, uint32_t n, const uint8_t* skip) {
    DbWriter wr;
    if (!wr.open(path)) return FCZ_E_ARG;
    std::vector<uint32_t> keys(n);
    std::vector<std::string> names(n);
    for (uint32_t c = 0; c < n; c++) { keys[c] = 100u + c; names[c] = "e" + std::to_string(c); }
    if (!wr.append("head", 4, 1, "head")) return FCZ_E_ARG;
    if (!wr.appendBatch(base, off, n, keys.data(), names.data(), skip)) return FCZ_E_ARG;
    if (!wr.append("tail", 4, 2, "tail")) return FCZ_E_ARG;
    return wr.close() ? FCZ_OK : FCZ_E_ARG;
}

extern "C" int fczgpu_pdb_title(const char* text, size_t len, const char* base_name, char* out, size_t cap) {
    const std::string t = pdbTitle(text, len, base_name);
    if (t.size() + 1 > cap) return -1;
    memcpy(out, t.c_str(), t.size() + 1);
    return (int)t.size();
}

// parsePdbUnits for bindings and tests: the units of one text, concatenated (res_off / atom_off [n_units + 1]).  Returns
// the parser's flag, or -1 when a capacity is too small; *n_units is set either way when the flag is 0.
extern "C" int fczgpu_parse_pdb_units(const char* text, size_t len, uint32_t* n_units, uint32_t* res_off, uint64_t* atom_off,
                                      uint8_t* res_type, float* bfactor, float* xyz, fcz_chain_meta* meta, uint32_t cap_units,
                                      uint32_t cap_res, uint64_t cap_atoms) {
    std::vector<CanonicalChain> units;
    const int flag = parsePdbUnits(text, len, "", units);
    if (flag) return flag;
    *n_units = (uint32_t)units.size();
    if (units.size() > cap_units) return -1;
    uint64_t r = 0, a = 0;
    for (size_t u = 0; u < units.size(); u++) {
        res_off[u] = (uint32_t)r; atom_off[u] = a;
        const CanonicalChain& c = units[u];
        if (r + c.res_type.size() > cap_res || a + c.xyz.size() / 3 > cap_atoms) return -1;
        memcpy(res_type + r, c.res_type.data(), c.res_type.size());
        memcpy(bfactor + r, c.bfactor.data(), 4 * c.bfactor.size());
        memcpy(xyz + 3 * a, c.xyz.data(), 4 * c.xyz.size());
        meta[u] = c.meta;
        r += c.res_type.size(); a += c.xyz.size() / 3;
    }
    res_off[units.size()] = (uint32_t)r; atom_off[units.size()] = a;
    return 0;
}

static void put_stats(const DbStats& s, double* o) {
    if (!o) return;
    o[0] = (double)s.entries; o[1] = (double)s.failed; o[2] = (double)s.residues; o[3] = (double)s.bytes_in; o[4] = (double)s.bytes_out;
    o[5] = s.seconds; o[6] = s.seconds_engine;
}
extern "C" int fczgpu_decompress_db(int device, const char* in_db, const char* out_db, int alt_order, double* stats7) {
    try {
        Engine eng(device);
        DbStats s;
        const int rc = decompressDb(eng, in_db, out_db, alt_order != 0, &s);
        put_stats(s, stats7);
        return rc;
    } catch (const std::exception&) {
        return FCZ_E_CUDA;
    }
}
extern "C" int fczgpu_compress_db(int device, const char* in_db, const char* out_db, int anchor_threshold, double* stats7) {
    try {
        Engine eng(device);
        DbStats s;
        const int rc = compressDb(eng, in_db, out_db, anchor_threshold, &s);
        put_stats(s, stats7);
        return rc;
    } catch (const std::exception&) {
        return FCZ_E_CUDA;
    }
}

// small hooks for bindings and tests
extern "C" float fczgpu_parse_float(const char* s, size_t n) { return parseFixedFloat(s, n); }
// read a database with DbReader and rewrite it with DbWriter (payloads kept, entries NUL-terminated, sorted by key)
extern "C" int fczgpu_db_copy(const char* in_db, const char* out_db) {
    DbReader rd;
    if (!rd.open(in_db)) return FCZ_E_ARG;
    DbWriter wr;
    if (!wr.open(out_db)) return FCZ_E_ARG;
    for (size_t i = 0; i < rd.size(); i++)
        if (!wr.append(rd.data(i), rd.payload(i), rd.key(i), rd.name(i))) return FCZ_E_ARG;
    return wr.close() ? (int)rd.size() : FCZ_E_ARG;
}

// ------------------------------------------------------------------ the reference's C-style database handles
// src/database_reader.h:11-27 / src/database_writer.h:12-15 over DbReader / DbWriter (C++ linkage, like the reference's).
namespace {
struct ReaderHandle {
    fczgpu::DbReader rd;
    std::vector<std::pair<std::string, uint32_t>> by_name;  // sorted by name
    bool lookup = false;
};
}  // namespace

void* make_reader(const char* data_name, const char* index_name, int32_t data_mode) {
    if (!data_name || !index_name) return nullptr;
    ReaderHandle* h = new ReaderHandle;
    if (!h->rd.open(data_name, index_name, (data_mode & 1) != 0)) { delete h; return nullptr; }
    h->lookup = (data_mode & (4 | 8)) != 0 && h->rd.hasLookup();
    if (h->lookup) {
        for (size_t i = 0; i < h->rd.size(); i++) h->by_name.emplace_back(h->rd.name(i), h->rd.key(i));
        std::sort(h->by_name.begin(), h->by_name.end());
    }
    return h;
}
void free_reader(void* r) { delete (ReaderHandle*)r; }
int64_t reader_get_size(void* r) { return r ? (int64_t)((ReaderHandle*)r)->rd.size() : -1; }
int64_t reader_get_id(void* r, uint32_t key) {
    if (!r) return -1;
    const fczgpu::DbReader& rd = ((ReaderHandle*)r)->rd;
    size_t lo = 0, hi = rd.size();
    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (rd.key(mid) < key) lo = mid + 1; else hi = mid; }
    return (lo < rd.size() && rd.key(lo) == key) ? (int64_t)lo : -1;
}
static bool in_range(void* r, int64_t id) { return r && id >= 0 && id < (int64_t)((ReaderHandle*)r)->rd.size(); }
const char* reader_get_data(void* r, int64_t id) { return in_range(r, id) && ((ReaderHandle*)r)->rd.hasData() ? ((ReaderHandle*)r)->rd.data((size_t)id) : nullptr; }
uint32_t reader_get_key(void* r, int64_t id) { return in_range(r, id) ? ((ReaderHandle*)r)->rd.key((size_t)id) : UINT32_MAX; }
int64_t reader_get_length(void* r, int64_t id) { return in_range(r, id) ? (int64_t)((ReaderHandle*)r)->rd.length((size_t)id) : -1; }
int64_t reader_get_offset(void* r, int64_t id) { return in_range(r, id) ? (int64_t)((ReaderHandle*)r)->rd.offset((size_t)id) : -1; }
uint32_t reader_lookup_entry(void* r, const char* name) {
    ReaderHandle* h = (ReaderHandle*)r;
    if (!h || !h->lookup || !name) return UINT32_MAX;
    const std::string n(name);
    auto it = std::lower_bound(h->by_name.begin(), h->by_name.end(), n, [](const std::pair<std::string, uint32_t>& a, const std::string& b) { return a.first < b; });
    return (it != h->by_name.end() && it->first == n) ? it->second : UINT32_MAX;
}
const char* reader_lookup_name_alloc(void* r, uint32_t key) {
    ReaderHandle* h = (ReaderHandle*)r;
    if (!h || !h->lookup) return "";
    const int64_t id = reader_get_id(r, key);
    if (id < 0) return "";
    return strdup(h->rd.name((size_t)id).c_str());
}
void* make_writer(const char* data_name, const char* index_name) {
    if (!data_name || !index_name) return nullptr;
    fczgpu::DbWriter* w = new fczgpu::DbWriter;
    if (!w->open(data_name, index_name)) { delete w; return nullptr; }
    return w;
}
void free_writer(void* w) {
    if (!w) return;
    ((fczgpu::DbWriter*)w)->close();
    delete (fczgpu::DbWriter*)w;
}
bool writer_append(void* w, const char* data, size_t length, uint32_t key, const char* name) {
    return w && ((fczgpu::DbWriter*)w)->appendRaw(data, length, key, name ? name : "");
}
