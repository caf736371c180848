/* Ahead-of-time instantiations, parity build: compiled with -fmad=false; the reference's operation order
 *  and draw injection (SURVEY.md §8c level L-A). */
#include "me_kernels.cuh"

static const MeAotEntry g_table[] = { ME_AOT_TABLE(true) };

extern "C" const MeAotEntry *me_aot_strict_table(int *n) {
    *n = (int)(sizeof(g_table) / sizeof(g_table[0]));
    return g_table;
}
