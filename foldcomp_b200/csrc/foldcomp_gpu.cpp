// foldcomp_b200/csrc/foldcomp_gpu.cpp -- see foldcomp_gpu.h.  Host-side data marshalling only.
#include "foldcomp_gpu.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>

#include "fcz_tables.h"

namespace fczgpu {

static int code_of(const std::string& name3) {  // getOneLetterCode + convertOneLetterCodeToInt: unknown -> UNK
    for (int c = 0; c < FCZ_NUM_CODES; c++)
        if (name3 == FCZ_NAME3[c]) return FCZ_NATOMS[c] ? c : FCZ_CODE_UNK;
    return FCZ_CODE_UNK;
}

CanonicalChain canonicalize(const AtomCoordinate* atoms, size_t n, const std::string& title) {
    CanonicalChain c;
    c.title = title;
    if (n == 0) return c;
    c.meta.n_atom = (uint16_t)n;
    c.meta.idx_residue = (uint16_t)atoms[0].residue_index;
    c.meta.idx_atom = (uint16_t)atoms[0].atom_index;
    c.meta.chain = atoms[0].chain.empty() ? ' ' : (uint8_t)atoms[0].chain[0];
    if (atoms[n - 1].atom == "OXT") {
        c.meta.has_oxt = 1;
        c.meta.oxt[0] = atoms[n - 1].coordinate.x;
        c.meta.oxt[1] = atoms[n - 1].coordinate.y;
        c.meta.oxt[2] = atoms[n - 1].coordinate.z;
    }
    size_t i = 0;
    while (i < n) {  // splitAtomByResidue (src/atom_coordinate.cpp:304-328): the last atom joins the current residue
        size_t j = i + 1;
        while (j < n && (atoms[j].residue_index == atoms[j - 1].residue_index || j == n - 1)) j++;
        const int code = code_of(atoms[i].residue);
        c.res_type.push_back((uint8_t)code);
        float bf = 0.f;
        for (int k = 0; k < FCZ_NATOMS[code]; k++) {
            const char* want = FCZ_ATOM_NAME[code][k];
            float x = 0, y = 0, z = 0;  // missing atom -> (0,0,0)
            for (size_t a = i; a < j; a++)
                if (atoms[a].atom == want) { x = atoms[a].coordinate.x; y = atoms[a].coordinate.y; z = atoms[a].coordinate.z; break; }
            c.xyz.push_back(x); c.xyz.push_back(y); c.xyz.push_back(z);
        }
        for (size_t a = i; a < j; a++)
            if (atoms[a].atom == "CA") { bf = atoms[a].tempFactor; break; }
        c.bfactor.push_back(bf);
        i = j;
    }
    return c;
}

void to_atoms(const CanonicalChain& c, bool alt_order, std::vector<AtomCoordinate>& atoms) {
    atoms.clear();
    const std::string chain(1, (char)c.meta.chain);
    int serial = c.meta.idx_atom;
    size_t a = 0;
    const size_t L = c.res_type.size();
    for (size_t r = 0; r < L; r++) {
        const int code = c.res_type[r];
        for (int k = 0; k < FCZ_NATOMS[code]; k++, a++) {
            AtomCoordinate at;
            // the engine already emitted coordinates in the requested order; names follow the same order
            at.atom = FCZ_ATOM_NAME[code][alt_order ? FCZ_ALT[code][k] : k];
            at.residue = FCZ_NAME3[code];
            at.chain = chain;
            at.atom_index = serial++;
            at.residue_index = c.meta.idx_residue + (int)r;
            at.coordinate.x = c.xyz[3 * a]; at.coordinate.y = c.xyz[3 * a + 1]; at.coordinate.z = c.xyz[3 * a + 2];
            at.tempFactor = c.bfactor[r];
            atoms.push_back(at);
        }
    }
    if (c.meta.has_oxt && L) {  // src/foldcomp.cpp:892-897
        AtomCoordinate at;
        at.atom = "OXT"; at.residue = FCZ_NAME3[c.res_type[L - 1]]; at.chain = chain;
        at.atom_index = serial++; at.residue_index = (int)L;  // Foldcomp::read builds OXT with residue_index = nResidue (958-961)
        at.coordinate.x = c.meta.oxt[0]; at.coordinate.y = c.meta.oxt[1]; at.coordinate.z = c.meta.oxt[2];
        at.tempFactor = c.bfactor[L - 1];
        atoms.push_back(at);
    }
}

Engine::Engine(int device) : e_(fcz_engine_create(device, nullptr)) {
    if (!e_) throw std::runtime_error("fcz_engine_create failed: no CUDA device (foldcomp_b200 has no CPU fallback)");
}
Engine::~Engine() { fcz_engine_destroy(e_); }

int FoldcompGpu::compressBatch(Engine& eng, const std::vector<CanonicalChain>& chains, int anchorThreshold,
                               std::vector<std::string>& blobs, std::vector<int>& status) {
    const uint32_t n = (uint32_t)chains.size();
    std::vector<uint32_t> res_off(n + 1, 0), title_off(n + 1, 0);
    std::vector<uint64_t> atom_off(n + 1, 0);
    for (uint32_t c = 0; c < n; c++) {
        res_off[c + 1] = res_off[c] + (uint32_t)chains[c].res_type.size();
        atom_off[c + 1] = atom_off[c] + chains[c].xyz.size() / 3;
        title_off[c + 1] = title_off[c] + (uint32_t)chains[c].title.size();
    }
    std::vector<uint8_t> res_type(res_off[n] + 1);
    std::vector<float> bfac(res_off[n] + 1), xyz(3 * atom_off[n] + 3);
    std::string titles;
    std::vector<fcz_chain_meta> meta(n + 1);
    for (uint32_t c = 0; c < n; c++) {
        memcpy(res_type.data() + res_off[c], chains[c].res_type.data(), chains[c].res_type.size());
        memcpy(bfac.data() + res_off[c], chains[c].bfactor.data(), 4 * chains[c].bfactor.size());
        memcpy(xyz.data() + 3 * atom_off[c], chains[c].xyz.data(), 4 * chains[c].xyz.size());
        titles += chains[c].title;
        meta[c] = chains[c].meta;
    }
    titles.push_back('\0');
    fcz_chain_batch in{};
    in.n_chains = n; in.mem = FCZ_MEM_HOST;
    in.res_off = res_off.data(); in.atom_off = atom_off.data(); in.title_off = title_off.data();
    in.res_type = res_type.data(); in.bfactor = bfac.data(); in.xyz = xyz.data();
    in.titles = &titles[0]; in.meta = meta.data();
    const uint64_t cap = fcz_encode_bound(n, res_off[n], atom_off[n], title_off[n], anchorThreshold) + 64;
    std::vector<uint8_t> bytes(cap);
    std::vector<uint64_t> blob_off(n + 1);
    std::vector<int32_t> st(n + 1);
    fcz_blob_batch out{};
    out.n_chains = n; out.mem = FCZ_MEM_HOST; out.blob_off = blob_off.data(); out.bytes = bytes.data();
    out.status = st.data(); out.bytes_cap = cap;
    fcz_opts o{anchorThreshold, 0, nullptr, 0};
    int rc = fcz_engine_set_opts(eng.get(), &o);
    if (rc) return rc;
    rc = fcz_encode_batch(eng.get(), &in, &out);
    if (rc) return rc;
    blobs.resize(n);
    status.assign(st.begin(), st.begin() + n);
    for (uint32_t c = 0; c < n; c++) blobs[c].assign((const char*)bytes.data() + blob_off[c], blob_off[c + 1] - blob_off[c]);
    return FCZ_OK;
}

int FoldcompGpu::decompressBatch(Engine& eng, const std::vector<std::string>& blobs, bool altOrder,
                                 std::vector<CanonicalChain>& chains, std::vector<int>& status) {
    const uint32_t n = (uint32_t)blobs.size();
    std::vector<uint64_t> blob_off(n + 1, 0);
    for (uint32_t c = 0; c < n; c++) blob_off[c + 1] = blob_off[c] + blobs[c].size();
    std::vector<uint8_t> bytes(blob_off[n] + 1);
    for (uint32_t c = 0; c < n; c++) memcpy(bytes.data() + blob_off[c], blobs[c].data(), blobs[c].size());
    fcz_blob_batch in{};
    in.n_chains = n; in.mem = FCZ_MEM_HOST; in.blob_off = blob_off.data(); in.bytes = bytes.data();
    std::vector<uint32_t> res_off(n + 1), title_off(n + 1);
    std::vector<uint64_t> atom_off(n + 1);
    std::vector<int32_t> st(n + 1);
    fcz_chain_batch out{};
    out.n_chains = n; out.mem = FCZ_MEM_HOST;
    out.res_off = res_off.data(); out.atom_off = atom_off.data(); out.title_off = title_off.data(); out.status = st.data();
    fcz_opts o{25, altOrder ? 1 : 0, nullptr, 0};
    int rc = fcz_engine_set_opts(eng.get(), &o);
    if (rc) return rc;
    fcz_sizes sz{};
    rc = fcz_decode_plan(eng.get(), &in, &out, &sz);
    if (rc) return rc;
    std::vector<uint8_t> res_type(sz.n_res + 1);
    std::vector<float> bfac(sz.n_res + 1), xyz(3 * sz.n_atoms + 3);
    std::vector<char> titles(sz.n_title_bytes + 1);
    std::vector<fcz_chain_meta> meta(n + 1);
    out.res_type = res_type.data(); out.bfactor = bfac.data(); out.xyz = xyz.data(); out.titles = titles.data();
    out.meta = meta.data(); out.res_cap = sz.n_res; out.atom_cap = sz.n_atoms; out.title_cap = sz.n_title_bytes;
    rc = fcz_decode_batch(eng.get(), &in, &out);
    if (rc) return rc;
    chains.assign(n, CanonicalChain());
    status.assign(st.begin(), st.begin() + n);
    for (uint32_t c = 0; c < n; c++) {
        if (st[c] != FCZ_OK) continue;
        CanonicalChain& ch = chains[c];
        ch.res_type.assign(res_type.begin() + res_off[c], res_type.begin() + res_off[c + 1]);
        ch.bfactor.assign(bfac.begin() + res_off[c], bfac.begin() + res_off[c + 1]);
        ch.xyz.assign(xyz.begin() + 3 * atom_off[c], xyz.begin() + 3 * atom_off[c + 1]);
        ch.title.assign(titles.data() + title_off[c], title_off[c + 1] - title_off[c]);
        ch.meta = meta[c];
    }
    return FCZ_OK;
}

int FoldcompGpu::compress(const std::vector<AtomCoordinate>& atoms) {
    std::vector<CanonicalChain> ch(1, canonicalize(atoms.data(), atoms.size(), strTitle));
    nAtom = (int)atoms.size();
    nResidue = (int)ch[0].res_type.size();
    std::vector<std::string> blobs;
    std::vector<int> st;
    int rc = compressBatch(eng_, ch, anchorThreshold, blobs, st);
    if (rc) return rc;
    if (st[0]) return st[0];
    blob_ = blobs[0];
    return 0;
}

int FoldcompGpu::writeStream(std::ostream& os) const {
    os.write(blob_.data(), (std::streamsize)blob_.size());
    return 0;
}

int FoldcompGpu::read(std::istream& is) {
    blob_.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
    if (blob_.size() < 4 || memcmp(blob_.data(), "FCMP", 4) != 0) return -1;
    return 0;
}

int FoldcompGpu::decompressToPdb(std::string& text) {
    uint64_t blob_off[2] = {0, blob_.size()}, text_off[2] = {0, 0};
    int32_t st[2] = {0, 0};
    fcz_blob_batch in{};
    in.n_chains = 1; in.mem = FCZ_MEM_HOST; in.blob_off = blob_off; in.bytes = (uint8_t*)&blob_[0];
    fcz_text_batch out{};
    out.n_chains = 1; out.mem = FCZ_MEM_HOST; out.text_off = text_off; out.status = st;
    fcz_opts o{anchorThreshold, useAltAtomOrder ? 1 : 0, nullptr, 0};
    int rc = fcz_engine_set_opts(eng_.get(), &o);
    if (rc) return rc;
    uint64_t total = 0;
    if ((rc = fcz_decode_to_pdb_plan(eng_.get(), &in, &out, &total))) return rc;
    if (st[0]) return st[0];
    text.assign(total, '\0');
    out.bytes = &text[0]; out.bytes_cap = total;
    return fcz_decode_to_pdb_batch(eng_.get(), &in, &out);
}

int FoldcompGpu::decompress(std::vector<AtomCoordinate>& atoms) {
    std::vector<CanonicalChain> ch;
    std::vector<int> st;
    int rc = decompressBatch(eng_, std::vector<std::string>(1, blob_), useAltAtomOrder, ch, st);
    if (rc) return rc;
    if (st[0]) return st[0];
    strTitle = ch[0].title;
    nResidue = (int)ch[0].res_type.size();
    to_atoms(ch[0], useAltAtomOrder, atoms);
    nAtom = (int)atoms.size();
    return 0;
}

int FoldcompGpu::write(const std::string& filename) const {  // Foldcomp::write, src/foldcomp.cpp:1111-1119
    std::ofstream out(filename, std::ios::out | std::ios::binary);
    if (!out) return -1;
    return writeStream(out);
}

// Foldcomp::writeTar (src/foldcomp.cpp:1122-1187) appends the entry to an open microtar archive; here the same record
// goes to a stream: the 512-byte header exactly as mtar_write_file_header + mtar_write_header build it
// (lib/microtar/microtar.c:186-220: name, mode 644, owner 0, size and mtime in unpadded octal, type '0', checksum
// "%06o" NUL ' '), the blob, zero padding to a 512-byte boundary.  writeTarEnd writes the two closing zero records
// (mtar_finalize).
int FoldcompGpu::writeTar(std::ostream& tar, const std::string& filename) const {
    char h[512];
    memset(h, 0, sizeof h);
    if (filename.size() >= 100) return -1;
    memcpy(h, filename.c_str(), filename.size());
    snprintf(h + 100, 8, "%o", 0644u);
    snprintf(h + 108, 8, "%o", 0u);
    snprintf(h + 124, 12, "%o", (unsigned)blob_.size());  // (116: the group field, left zero as microtar leaves it)
    snprintf(h + 136, 12, "%o", 0u);
    h[148 + 8] = '0';  // type, right after the 8-byte checksum field at offset 148
    unsigned sum = 256;
    for (int i = 0; i < 148; i++) sum += (unsigned char)h[i];
    for (int i = 156; i < 512; i++) sum += (unsigned char)h[i];
    snprintf(h + 148, 8, "%06o", sum);
    h[155] = ' ';
    tar.write(h, 512);
    tar.write(blob_.data(), (std::streamsize)blob_.size());
    static const char zeros[512] = {0};
    const size_t pad = (512 - blob_.size() % 512) % 512;
    tar.write(zeros, (std::streamsize)pad);
    return tar ? 0 : -1;
}
int FoldcompGpu::writeTarEnd(std::ostream& tar) {
    static const char zeros[1024] = {0};
    tar.write(zeros, 1024);
    return tar ? 0 : -1;
}

int FoldcompGpu::extract(std::string& data, int type, int digits) {  // Foldcomp::extract, src/foldcomp.cpp:1260-1336
    uint64_t blob_off[2] = {0, blob_.size()}, text_off[2] = {0, 0}, total = 0;
    fcz_blob_batch in{};
    in.n_chains = 1; in.mem = FCZ_MEM_HOST; in.blob_off = blob_off; in.bytes = (uint8_t*)&blob_[0];
    std::string out(blob_.size() + 64, '\0');  // <= 6 characters per residue, >= 9 blob bytes per residue
    fcz_text_batch tb{};
    tb.n_chains = 1; tb.mem = FCZ_MEM_HOST; tb.text_off = text_off; tb.bytes = &out[0]; tb.bytes_cap = out.size();
    const int rc = fcz_extract_batch(eng_.get(), &in, type, digits, &tb, &total);
    if (rc) return rc;
    data.append(out.data(), (size_t)total);
    return 0;
}

int FoldcompGpu::checkValidity() {  // Foldcomp::checkValidity, src/foldcomp.cpp:1492-1532: the ValidityError class 0..6
    uint64_t blob_off[2] = {0, blob_.size()};
    fcz_blob_batch in{};
    in.n_chains = 1; in.mem = FCZ_MEM_HOST; in.blob_off = blob_off; in.bytes = (uint8_t*)&blob_[0];
    int32_t rs = 0, v = 0;
    const int rc = fcz_check_batch(eng_.get(), &in, &rs, &v);
    return rc ? rc : v;
}

}  // namespace fczgpu
