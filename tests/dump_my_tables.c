/* tests/dump_my_tables.c -- prints foldcomp_b200/csrc/fcz_tables.h in the format of
 * oracle/dump_tables.cpp so the two can be diffed (tests/test_tables.py). */
#include <stdio.h>
#include "../foldcomp_b200/csrc/fcz_tables.h"
int main(void) {
    for (int c = 0; c < 20; c++) {
        int n = FCZ_NATOMS[c];
        printf("%d %s %d\n", c, FCZ_NAME3[c], n);
        for (int k = 3; k < n; k++) {
            unsigned p = FCZ_PRED[c][k];
            printf("%d %s %d %d %d %a %a\n", k, FCZ_ATOM_NAME[c][k], p & 15, (p >> 4) & 15, (p >> 8) & 15,
                   FCZ_BLEN[c][k], FCZ_BANG[c][k]);
        }
        printf("alt");
        for (int k = 0; k < n; k++) printf(" %d", FCZ_ALT[c][k]);
        printf("\n");
    }
    return 0;
}
