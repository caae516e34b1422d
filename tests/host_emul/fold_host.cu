// TEST INFRASTRUCTURE ONLY: launch wrappers for the per-tap dy-fold kernels of the small-map dgrad path (csrc/fold_kernels.cuh;
// in the product they are internal to bg_conv2d_dgrad).  Compiled for the host with -DBG_HOST_EMULATION by
// tests/test_fold_kernels_host_emul.py.
#include "../../biggan_tensorflow_b200/csrc/bg_common.cuh"

#include "../../biggan_tensorflow_b200/csrc/fold_kernels.cuh"

extern "C" int test_reflect_dyfold(const void* dy, void* dyf, int B, int H, int W, int C, int use4) {
  if (use4) {
    if (H != 4 || W != 4) return -1;
    long blocks = bg_cdiv((long)B * (C / 8), 128);
    BG_LAUNCH(reflect_dyfold4_kernel, (int)blocks, 128, nullptr, (const bf16*)dy, (bf16*)dyf, B, C);
  } else {
    long blocks = bg_cdiv((long)B * 9 * H * W * (C / 8), 256);
    if (blocks > 7) blocks = 7;                    // fewer blocks than work: the grid-stride loop is exercised
    BG_LAUNCH(reflect_dyfold_kernel, (int)blocks, 256, nullptr, (const bf16*)dy, (bf16*)dyf, B, H, W, C);
  }
  return 0;
}
