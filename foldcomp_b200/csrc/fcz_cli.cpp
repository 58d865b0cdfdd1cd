// foldcomp_b200/csrc/fcz_cli.cpp -- minimal `compress` / `decompress` front end over FoldcompGpu, shaped
// like the per-entry lambdas of the reference CLI (src/main.cpp:438-536, 612-689): one single-chain PDB
// file <-> one .fcz file.  ATOM records are read with the fixed columns the reference's CPython module
// uses (foldcomp/foldcomp.cxx:253-293) and written like writeAtomCoordinatesToPDB
// (src/atom_coordinate.cpp:220-291).  Host text I/O only; the codec runs on the GPU.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "fcz_db.h"
#include "foldcomp_gpu.h"

using namespace fczgpu;

static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

static int read_pdb(const std::string& path, std::vector<AtomCoordinate>& atoms) {
    std::ifstream f(path);
    if (!f) return 1;
    std::string line, chain;
    while (std::getline(f, line)) {
        if (line.compare(0, 4, "ATOM") != 0 || line.size() < 66) continue;
        if (chain.empty()) chain = line.substr(21, 1);
        if (line.substr(21, 1) != chain) return 2;  // multiple chains (foldcomp.cxx:266-268)
        AtomCoordinate a;
        a.atom = trim(line.substr(12, 4)); a.residue = trim(line.substr(17, 3)); a.chain = chain;
        a.atom_index = std::stoi(line.substr(6, 5)); a.residue_index = std::stoi(line.substr(22, 4));
        a.coordinate.x = std::stof(line.substr(30, 8)); a.coordinate.y = std::stof(line.substr(38, 8));
        a.coordinate.z = std::stof(line.substr(46, 8));
        a.occupancy = std::stof(line.substr(54, 6)); a.tempFactor = std::stof(line.substr(60, 6));
        if (!atoms.empty() && atoms.back().atom == a.atom) continue;  // removeAlternativePosition
        atoms.push_back(a);
    }
    return atoms.empty() ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: fcz_cli compress [-b N] in.pdb out.fcz | decompress [-a] in.fcz out.pdb\n"
                        "       fcz_cli compress-db [-b N] in_pdb_db out_fcz_db | decompress-db [-a] in_fcz_db out_pdb_db\n"
                        "       fcz_cli compress-tar [-b N] in.pdb out.tar | extract-plddt | extract-fasta | check  in.fcz out.txt\n");
        return 2;
    }
    const std::string mode = argv[1];
    int ai = 2, b = 25;
    bool alt = false;
    while (ai < argc && argv[ai][0] == '-') {
        if (!strcmp(argv[ai], "-b") && ai + 1 < argc) { b = atoi(argv[ai + 1]); ai += 2; }
        else if (!strcmp(argv[ai], "-a")) { alt = true; ai++; }
        else break;
    }
    if (ai + 2 > argc) return 2;
    const std::string in = argv[ai], out = argv[ai + 1];
    try {
        const auto t_start = std::chrono::steady_clock::now();
        Engine eng(0);
        FoldcompGpu comp(eng);
        const double init_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        if (mode == "compress-db" || mode == "decompress-db") {  // whole foldcomp databases, batched over the engine (fcz_db.h)
            DbStats st;
            const int rc = mode == "compress-db" ? compressDb(eng, in, out, b, &st) : decompressDb(eng, in, out, alt, &st);
            if (rc) { fprintf(stderr, "[Error] %s: %s\n", mode.c_str(), fcz_strerror(rc)); return 1; }
            fprintf(stderr, "%zu entries (%zu failed), %llu residues, %.3f s total, %.3f s in the engine, %.3f s CUDA context + engine start\n",
                    st.entries, st.failed, (unsigned long long)st.residues, st.seconds, st.seconds_engine, init_s);
            fflush(nullptr);
            _exit(0);  // the outputs are closed: skip unmapping the inputs and tearing down the CUDA context
        }
        if (mode == "compress") {
            // one PDB file like the reference CLI's single-file mode (src/main.cpp:438-530): the title by its rule (HEADER id,
            // TITLE records, else the OUTPUT's name without extension, 451-467); one .fcz per chain and fragment of continuous
            // numbering, named <out><chain>_<fragment>.fcz when there are several (494-509)
            std::ifstream f(in, std::ios::binary);
            if (!f) { fprintf(stderr, "[Error] cannot read %s\n", in.c_str()); return 1; }
            std::stringstream ss;
            ss << f.rdbuf();
            const std::string text = ss.str();
            const std::string in_base = in.substr(in.find_last_of('/') == std::string::npos ? 0 : in.find_last_of('/') + 1);
            const std::string in_stem = in_base.substr(0, in_base.find_last_of('.'));
            const size_t dot = out.find_last_of('.'), slash = out.find_last_of('/');
            const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
            const std::string out_stem = has_ext ? out.substr(0, dot) : out, out_ext = has_ext ? out.substr(dot) : std::string(".fcz");
            std::string title = pdbTitle(text.data(), text.size(), in_base);
            if (title == in_stem) title = out_stem;  // the reader fell back to the file name: single-file mode takes getFileParts(output).first,
                                                     // i.e. the output path as given without its extension (src/utility.cpp:118-126)
            std::vector<CanonicalChain> units;
            std::vector<UnitLabel> labels;
            int rc = parsePdbUnits(text.data(), text.size(), title, units, &labels);
            if (rc) { fprintf(stderr, "[Error] %s\n", rc == 1 ? "no atoms found" : "malformed ATOM record"); return 1; }
            std::vector<std::string> blobs;
            std::vector<int> st;
            if ((rc = FoldcompGpu::compressBatch(eng, units, b, blobs, st))) { fprintf(stderr, "[Error] compress: %s\n", fcz_strerror(rc)); return 1; }
            for (size_t u = 0; u < units.size(); u++) {
                if (st[u] != FCZ_OK) { fprintf(stderr, "[Error] compress: %s\n", fcz_strerror(st[u])); return 1; }
                std::string name = out_stem;
                if (labels[u].n_chains > 1) name += labels[u].chain;
                if (labels[u].n_frags > 1) name += "_" + std::to_string(labels[u].frag);
                std::ofstream os(name + out_ext, std::ios::binary);
                os.write(blobs[u].data(), (std::streamsize)blobs[u].size());
            }
        } else if (mode == "decompress") {
            std::ifstream is(in, std::ios::binary);
            if (!is || comp.read(is) != 0) { fprintf(stderr, "[Error] not an FCZ file\n"); return 1; }
            comp.useAltAtomOrder = alt;
            std::string text;
            int rc = comp.decompressToPdb(text);  // decode + writeAtomCoordinatesToPDB-identical text, both on the GPU
            if (rc) { fprintf(stderr, "[Error] decompress: %s\n", fcz_strerror(rc)); return 1; }
            std::ofstream os(out, std::ios::binary);
            os.write(text.data(), (std::streamsize)text.size());
        } else if (mode == "compress-tar") {  // one-member archive: Foldcomp::writeTar (src/main.cpp:518-523)
            std::vector<AtomCoordinate> atoms;
            int rc = read_pdb(in, atoms);
            if (rc) { fprintf(stderr, "[Error] %s\n", rc == 2 ? "multiple chains" : "no atoms found"); return 1; }
            std::string base = in.substr(in.find_last_of('/') == std::string::npos ? 0 : in.find_last_of('/') + 1);
            base = base.substr(0, base.find_last_of('.'));
            comp.strTitle = base;
            comp.anchorThreshold = b;
            if ((rc = comp.compress(atoms))) { fprintf(stderr, "[Error] compress: %s\n", fcz_strerror(rc)); return 1; }
            std::ofstream os(out, std::ios::binary);
            if (comp.writeTar(os, base + ".fcz") || FoldcompGpu::writeTarEnd(os)) return 1;
        } else if (mode == "extract-plddt" || mode == "extract-fasta" || mode == "check") {
            std::ifstream is(in, std::ios::binary);
            if (!is || comp.read(is) != 0) { fprintf(stderr, "[Error] File is not a valid fcz file\n"); return 1; }
            std::ofstream os(out, std::ios::binary);
            if (mode == "check") {
                const int v = comp.checkValidity();  // ValidityError class (src/foldcomp.h:59-67)
                os << v << "\n";
            } else {
                std::string data;
                const int rc = comp.extract(data, mode == "extract-fasta" ? 1 : 0, 1);
                if (rc) { fprintf(stderr, "[Error] extract: %s\n", fcz_strerror(rc)); return 1; }
                os << ">" << in << "\n" << data << "\n";  // Foldcomp::writeFASTALike (src/foldcomp.cpp:1223-1231)
            }
        } else return 2;
    } catch (const std::exception& ex) {
        fprintf(stderr, "[Error] %s\n", ex.what());
        return 1;
    }
    return 0;
}
