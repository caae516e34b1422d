// TEST INFRASTRUCTURE ONLY: host build of csrc/next_ops.cu (the real kernel source under the emulation prelude) plus
// plain-loop stand-ins for the three already-hardware-verified entry points the Python glue of the 8(f) ops calls
// around them (bg_add, bg_axpby, bg_bgemm -- fp32 only).  tests/test_next_ops_glue_host.py loads this in place of
// libbiggan_b200.so to execute the glue code (argument order, tape, gradient accumulation) on CPU tensors.
#define BG_HOST_EMULATION 1
#include "../../biggan_tensorflow_b200/csrc/next_ops.cu"

extern "C" const char* bg_last_error(void) { return bg_emul_error; }

extern "C" int bg_add(const void* a, const void* b, void* out, long n, int dtype, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (long i = 0; i < n; ++i) ((float*)out)[i] = ((const float*)a)[i] + ((const float*)b)[i];
  return BG_OK;
}
extern "C" int bg_axpby(const void* x, const void* y, const float* alpha_dev, float alpha, float beta, void* out, long n, int dtype, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  const float a = alpha_dev ? *alpha_dev : alpha;
  for (long i = 0; i < n; ++i) ((float*)out)[i] = a * ((const float*)x)[i] + (y ? beta * ((const float*)y)[i] : 0.f);
  return BG_OK;
}
extern "C" int bg_bgemm(const void* A, const void* Bm, void* C, int batch, int M, int N, int K, long a_rs, long a_cs, long a_bs,
                        long b_rs, long b_cs, long b_bs, long c_rs, long c_bs, float alpha, int accumulate, int dtype, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (int b = 0; b < batch; ++b)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k)
          acc += (double)((const float*)A)[b * a_bs + m * a_rs + k * a_cs] * ((const float*)Bm)[b * b_bs + k * b_rs + n * b_cs];
        float* c = (float*)C + b * c_bs + m * c_rs + n;
        *c = (float)(alpha * acc) + (accumulate ? *c : 0.f);
      }
  return BG_OK;
}
