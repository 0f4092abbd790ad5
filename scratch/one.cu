#include "../woltka_b200/csrc/wk_seg.cuh"
using namespace wk;
void* f() { return (void*)classify_seg_kernel<WK_KIND_RANK, FX_FRAC, 512, false, false>; }
