#include "../woltka_b200/csrc/wk_sweep.cuh"
using namespace wk;
void* f() { return (void*)classify_sweep_kernel<true, SINK_DIRECT, true, 1024>; }
