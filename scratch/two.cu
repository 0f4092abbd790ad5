#include "../woltka_b200/csrc/wk_ordfuse.cuh"
using namespace wk;
void* f() { return (void*)ordinal_fused_kernel<FX_FRAC, false>; }
void* g() { return (void*)ordinal_listed_kernel; }
