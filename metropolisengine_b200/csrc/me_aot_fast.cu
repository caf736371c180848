/* Ahead-of-time instantiations, throughput build: FMA contraction on, reciprocal forms of the divisions. */
#include "me_kernels.cuh"

static const MeAotEntry g_table[] = { ME_AOT_TABLE(false) };

extern "C" const MeAotEntry *me_aot_fast_table(int *n) {
    *n = (int)(sizeof(g_table) / sizeof(g_table[0]));
    return g_table;
}
