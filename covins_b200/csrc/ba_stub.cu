#include "cvb_internal.cuh"
extern "C" void cvb_ba_free(cvb_ctx*) {}
