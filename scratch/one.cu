#include "../woltka_b200/csrc/wk_sweep.cuh"
using namespace wk;
void* f() { return (void*)classify_fast_kernel<WK_KIND_RANK, FX_FRAC, 13, false>; }
