// sdv_ba.cu — sliding-window back-end (placeholder until the BA kernels land; keeps the link closed)
#include "sdv_ctx.cuh"
namespace sdv { void ba_destroy(sdv_ctx*) {} }
