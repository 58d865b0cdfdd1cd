// oracle/dump_tables.cpp -- TEST INFRASTRUCTURE.
// Prints the reference's per-residue build table (src/amino_acid.h:69-406) in a flat text form:
//   <code> <NAME> <natoms>
//   <slot> <atom> <p0slot> <p1slot> <p2slot> <bondLength hexfloat> <bondAngle hexfloat>   (slot >= 3)
//   alt <slot order ...>
// tests/test_tables.py compares this (committed as tests/golden/aa_table_dump.txt) with the
// integer-slot table in foldcomp_b200/csrc/fcz_tables.h.
#include "amino_acid.h"

#include <cstdio>
#include <string>
#include <vector>

static const char* kNames[20] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
                                 "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"};

static int slot_of(const std::vector<std::string>& atoms, const std::string& n) {
    for (size_t i = 0; i < atoms.size(); i++)
        if (atoms[i] == n) return (int)i;
    return -1;
}

int main() {
    std::map<std::string, AminoAcid> aas = AminoAcid::AminoAcids();
    for (int c = 0; c < 20; c++) {
        const AminoAcid& aa = aas.at(kNames[c]);
        printf("%d %s %zu\n", c, kNames[c], aa.atoms.size());
        for (size_t k = 3; k < aa.atoms.size(); k++) {
            const std::string& x = aa.atoms[k];
            const std::vector<std::string>& p = aa.sideChain.at(x);
            float len = aa.bondLengths.at(p[2] + "_" + x);
            float ang = aa.bondAngles.at(p[1] + "_" + p[2] + "_" + x);
            printf("%zu %s %d %d %d %a %a\n", k, x.c_str(), slot_of(aa.atoms, p[0]),
                   slot_of(aa.atoms, p[1]), slot_of(aa.atoms, p[2]), len, ang);
        }
        printf("alt");
        for (const std::string& a : aa.altAtoms) printf(" %d", slot_of(aa.atoms, a));
        printf("\n");
    }
    return 0;
}
