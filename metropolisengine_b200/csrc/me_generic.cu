/* me_generic.cu — ahead-of-time instantiation of the runtime-shape kernels (me_generic.cuh) with the built-in functors;
 * compiled with -fmad=false (the parity instantiation for these shapes). */
#include "../../include/me_b200.h"
#include "me_generic.cuh"

extern "C" void me_generic_kernels(const void **run_measure, const void **init, const void **propose, const void **accept,
                                   const void **energy) {
    *run_measure = (const void *)&gk_run;
    *init = (const void *)&gk_init;
    *propose = (const void *)&gk_propose;
    *accept = (const void *)&gk_accept;
    *energy = (const void *)&gk_energy;
}
