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

// ---- loop stand-ins for the (hardware-verified) batch-norm entry points of norm.cu, fp32 only, same argument meaning:
//      sums = [sum x | sum x^2] per channel; mean / biased var + moving-average update; y = act(x inv + shift);
//      dgamma[b,c] = sum g xhat, dbeta[b,c] = sum g (g = dy masked by the ReLU); s = [sum_b gamma dbeta | sum_b gamma dgamma];
//      dx = rstd (gamma g - s1/count - xhat s2/count).
extern "C" size_t bg_bn_workspace_bytes(int, long, int) { return 256; }
extern "C" int bg_bn_stats(const void* x, double* sums, int B, long HW, int C, int dtype, void*, size_t, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (int c = 0; c < 2 * C; ++c) sums[c] = 0.0;
  const float* p = (const float*)x;
  for (long i = 0; i < (long)B * HW; ++i)
    for (int c = 0; c < C; ++c) { const double v = p[i * C + c]; sums[c] += v; sums[C + c] += v * v; }
  return BG_OK;
}
extern "C" int bg_bn_finalize(const double* sums, double count, float* mean, float* var, float* run_mean, float* run_var, float momentum,
                              float var_correction, int C, void*) {
  for (int c = 0; c < C; ++c) {
    const double m = sums[c] / count;
    double v = sums[C + c] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean[c] = (float)m; var[c] = (float)v;
    if (run_mean) run_mean[c] = run_mean[c] * momentum + (float)m * (1.f - momentum);
    if (run_var) run_var[c] = run_var[c] * momentum + (float)v * var_correction * (1.f - momentum);
  }
  return BG_OK;
}
static inline float bn_act(float v, int act) { return act == BG_ACT_RELU ? (v > 0.f ? v : 0.f) : v; }
extern "C" int bg_bn_apply(const void* x, const float* mean, const float* var, const float* gamma, const float* beta, int gb_bstride, void* y,
                           int act, float eps, int B, long HW, int C, int dtype, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (int b = 0; b < B; ++b)
    for (long r = 0; r < HW; ++r)
      for (int c = 0; c < C; ++c) {
        const float inv = 1.f / std::sqrt(var[c] + eps) * gamma[(long)b * gb_bstride + c];
        const long i = ((long)b * HW + r) * C + c;
        ((float*)y)[i] = bn_act(((const float*)x)[i] * inv + (beta[(long)b * gb_bstride + c] - mean[c] * inv), act);
      }
  return BG_OK;
}
extern "C" int bg_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* var, const float* gamma, const float* beta,
                                int gb_bstride, float* dgamma, float* dbeta, int act, float eps, int B, long HW, int C, int dtype, void*,
                                size_t, void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float rs = 1.f / std::sqrt(var[c] + eps), gm = gamma[(long)b * gb_bstride + c], inv = rs * gm;
      const float shift = beta[(long)b * gb_bstride + c] - mean[c] * inv;
      double sg = 0, sgx = 0;
      for (long r = 0; r < HW; ++r) {
        const long i = ((long)b * HW + r) * C + c;
        const float xv = ((const float*)x)[i];
        float g = ((const float*)dy)[i];
        if (act == BG_ACT_RELU && !(xv * inv + shift > 0.f)) g = 0.f;
        sg += g; sgx += (double)g * (xv - mean[c]) * rs;
      }
      dgamma[(long)b * C + c] = (float)sgx; dbeta[(long)b * C + c] = (float)sg;
    }
  return BG_OK;
}
extern "C" int bg_bn_bwd_sums(const float* dgamma, const float* dbeta, const float* gamma, int gb_bstride, double* s, int B, int C, void*) {
  for (int c = 0; c < C; ++c) {
    double a = 0, q = 0;
    for (int b = 0; b < B; ++b) { const double gm = gamma[(long)b * gb_bstride + c]; a += gm * dbeta[(long)b * C + c]; q += gm * dgamma[(long)b * C + c]; }
    s[c] = a; s[C + c] = q;
  }
  return BG_OK;
}
extern "C" int bg_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* var, const float* gamma, const float* beta,
                               int gb_bstride, const double* s, double count, void* dx, int act, float eps, int B, long HW, int C, int dtype,
                               void*) {
  if (dtype != BG_F32) { bg_set_error("host stand-in: fp32 only"); return BG_ERR_UNSUPPORTED; }
  for (int b = 0; b < B; ++b)
    for (long r = 0; r < HW; ++r)
      for (int c = 0; c < C; ++c) {
        const float rs = 1.f / std::sqrt(var[c] + eps), gm = gamma[(long)b * gb_bstride + c], inv = rs * gm;
        const float shift = beta[(long)b * gb_bstride + c] - mean[c] * inv;
        const long i = ((long)b * HW + r) * C + c;
        const float xv = ((const float*)x)[i];
        float g = ((const float*)dy)[i];
        if (act == BG_ACT_RELU && !(xv * inv + shift > 0.f)) g = 0.f;
        ((float*)dx)[i] = rs * (gm * g - (float)(s[c] / count) - (xv - mean[c]) * rs * (float)(s[C + c] / count));
      }
  return BG_OK;
}
