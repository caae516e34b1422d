// placeholder until the tcgen05 path lands
#include "bg_common.cuh"
int bg_tc_conv_fwd_type(const BgGeom& g, const bg_conv_desc* d, int op, const void* in, const void* w_k_major,
                        const float* bias, const void* residual, void* out, int act, cudaStream_t st, int* taken) { *taken = 0; return BG_OK; }
int bg_tc_conv_wgrad(const BgGeom& g, const bg_conv_desc* d, const void* in, const void* og, float* dw,
                     void* ws, size_t ws_bytes, cudaStream_t st, int* taken) { *taken = 0; return BG_OK; }
size_t bg_tc_conv_wgrad_ws(const BgGeom& g, const bg_conv_desc* d) { return 0; }
extern "C" int bg_has_tensor_core_path(void) { return 0; }
