// Hinge GAN losses (ops.py:753-797, :799-840, :847-848), logistic class loss (utils.py:366-369),
// fused Adam + EMA (tf.train.AdamOptimizer + MovingAverageOptimizer, BigGAN.py:923-926).
#include "bg_common.cuh"

// loss = sum relu(1-real)*inv + sum relu(1+fake)*inv ; flood: |L-f|+f.  Gradients written directly.
template <typename T>
__global__ void __launch_bounds__(256)
hinge_d_kernel(const T* __restrict__ real, const T* __restrict__ fake, int n_real, int n_fake, float inv_count, float flood,
               float* __restrict__ loss, T* __restrict__ dreal, T* __restrict__ dfake) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_real; i += blockDim.x) s += fmaxf(1.f - bg_ld(real, i), 0.f);
  for (int i = threadIdx.x; i < n_fake; i += blockDim.x) s += fmaxf(1.f + bg_ld(fake, i), 0.f);
  float L = bg_block_sum(s, sh) * inv_count;
  float sgn = 1.f;
  if (flood != 0.f) {
    const float d = L - flood;
    sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    L = fabsf(d) + flood;
  }
  if (threadIdx.x == 0) loss[0] = L;
  for (int i = threadIdx.x; i < n_real; i += blockDim.x)
    bg_st(dreal, i, (1.f - bg_ld(real, i) > 0.f) ? -sgn * inv_count : 0.f);
  for (int i = threadIdx.x; i < n_fake; i += blockDim.x)
    bg_st(dfake, i, (1.f + bg_ld(fake, i) > 0.f) ? sgn * inv_count : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
hinge_g_kernel(const T* __restrict__ fake, int n, float inv_count, float flood, float* __restrict__ loss, T* __restrict__ dfake) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += bg_ld(fake, i);
  float L = -bg_block_sum(s, sh) * inv_count;
  float sgn = 1.f;
  if (flood != 0.f) {
    const float d = L - flood;
    sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    L = fabsf(d) + flood;
  }
  if (threadIdx.x == 0) loss[0] = L;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bg_st(dfake, i, -sgn * inv_count);
}

extern "C" int bg_hinge_d_fwd_bwd(const void* real, const void* fake, int n_real, int n_fake, float inv_count, float flood,
                                  float* loss, void* dreal, void* dfake, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, hinge_d_kernel<T><<<1, 256, 0, (cudaStream_t)stream>>>((const T*)real, (const T*)fake, n_real, n_fake, inv_count,
                                                                                 flood, loss, (T*)dreal, (T*)dfake));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_hinge_g_fwd_bwd(const void* fake, int n, float inv_count, float flood, float* loss, void* dfake, int dtype,
                                  void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, hinge_g_kernel<T><<<1, 256, 0, (cudaStream_t)stream>>>((const T*)fake, n, inv_count, flood, loss, (T*)dfake));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// loss (+)= weight * inv_count * sum_{b,l} w_l * xent(labels, logits);  dlogits = weight*inv_count*w_l*(sigmoid(x)-z)
template <typename T>
__global__ void __launch_bounds__(1024)
logistic_cls_kernel(const T* __restrict__ logits, const float* __restrict__ labels, const float* __restrict__ cls_w, int B,
                    int n_labels, float weight, float inv_count, float* __restrict__ loss, T* __restrict__ dlogits,
                    int accumulate_loss, float* __restrict__ part) {
  __shared__ float sh[33];
  const long n = (long)B * n_labels;
  float s = 0.f;
  // blocks take interleaved 1024-element slabs; with `part` every block leaves its partial sum and logistic_sum_kernel adds
  // them in block order (deterministic), without it the launch is a single block that finishes the loss itself
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float x = bg_ld(logits, i), z = labels[i];
    const float w = cls_w ? cls_w[i % n_labels] : 1.f;
    s += w * (fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x))));
    const float sig = 1.f / (1.f + expf(-x));
    bg_st(dlogits, i, weight * inv_count * w * (sig - z));
  }
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) {
    if (part) part[blockIdx.x] = s;
    else loss[0] = (accumulate_loss ? loss[0] : 0.f) + weight * inv_count * s;
  }
}
__global__ void logistic_sum_kernel(const float* __restrict__ part, int np, float scale, float* __restrict__ loss, int accumulate_loss) {
  float s = 0.f;
  for (int i = 0; i < np; ++i) s += part[i];
  loss[0] = (accumulate_loss ? loss[0] : 0.f) + scale * s;
}

extern "C" size_t bg_logistic_cls_workspace_bytes(int B, int n_labels) { (void)B; (void)n_labels; return 256 * sizeof(float); }

extern "C" int bg_logistic_cls_fwd_bwd(const void* logits, const float* labels, const float* cls_w, int B, int n_labels,
                                       float weight, float inv_count, float* loss, void* dlogits, int accumulate_loss,
                                       int dtype, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long n = (long)B * n_labels;
  int nb = (int)min(bg_cdiv(n, 4096), 148L);
  if (!ws || ws_bytes < 256 * sizeof(float) || nb <= 1) {
    BG_DISPATCH_DTYPE(dtype, T, logistic_cls_kernel<T><<<1, 1024, 0, st>>>((const T*)logits, labels, cls_w, B, n_labels, weight, inv_count, loss,
                                                                        (T*)dlogits, accumulate_loss, nullptr));
    BG_LAUNCH_CHECK();
    return BG_OK;
  }
  BG_DISPATCH_DTYPE(dtype, T, logistic_cls_kernel<T><<<nb, 1024, 0, st>>>((const T*)logits, labels, cls_w, B, n_labels, weight, inv_count, loss,
                                                                       (T*)dlogits, accumulate_loss, (float*)ws));
  BG_LAUNCH_CHECK();
  logistic_sum_kernel<<<1, 1, 0, st>>>((const float*)ws, nb, weight * inv_count, loss, accumulate_loss);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p -= lr_t * m / (sqrt(v) + eps) ; ema = d*ema + (1-d)*p
__global__ void __launch_bounds__(256)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ ema, long n, float lr_host, const float* __restrict__ lr_dev, float b1, float b2, float eps, float d,
                float gscale) {
  const float lr_t = lr_dev ? lr_dev[0] : lr_host;       // device scalar: the step size can change under CUDA-graph replay
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<float4*>(m + i);
      float4 v4 = *reinterpret_cast<float4*>(v + i);
      float4 p4 = *reinterpret_cast<float4*>(p + i);
      float ge[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
      float me[4] = {m4.x, m4.y, m4.z, m4.w}, ve[4] = {v4.x, v4.y, v4.z, v4.w}, pe[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        me[e] = b1 * me[e] + (1.f - b1) * ge[e];
        ve[e] = b2 * ve[e] + (1.f - b2) * ge[e] * ge[e];
        pe[e] = pe[e] - lr_t * me[e] / (sqrtf(ve[e]) + eps);
      }
      *reinterpret_cast<float4*>(m + i) = make_float4(me[0], me[1], me[2], me[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(ve[0], ve[1], ve[2], ve[3]);
      *reinterpret_cast<float4*>(p + i) = make_float4(pe[0], pe[1], pe[2], pe[3]);
      if (ema) {
        float4 e4 = *reinterpret_cast<float4*>(ema + i);
        e4.x = d * e4.x + (1.f - d) * pe[0]; e4.y = d * e4.y + (1.f - d) * pe[1];
        e4.z = d * e4.z + (1.f - d) * pe[2]; e4.w = d * e4.w + (1.f - d) * pe[3];
        *reinterpret_cast<float4*>(ema + i) = e4;
      }
    } else {
      for (long j = i; j < n; ++j) {
        const float gj = g[j] * gscale;
        const float mj = b1 * m[j] + (1.f - b1) * gj;
        const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
        const float pj = p[j] - lr_t * mj / (sqrtf(vj) + eps);
        m[j] = mj; v[j] = vj; p[j] = pj;
        if (ema) ema[j] = d * ema[j] + (1.f - d) * pj;
      }
    }
  }
}

extern "C" int bg_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long n, float lr_t, const float* lr_t_dev,
                                float beta1, float beta2, float eps, float ema_decay, float grad_scale, void* stream) {
  if (n <= 0) return BG_OK;
  long blocks = bg_cdiv(n, 1024);
  const long mx = (long)bg_num_sms() * 8;
  if (blocks > mx) blocks = mx;
  adam_ema_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, n, lr_t, lr_t_dev, beta1, beta2, eps, ema_decay, grad_scale);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// state = {t, lr_t}: t += 1; lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)  (tf.train.AdamOptimizer); keeping the step
// counter on the device makes the whole training step replayable as one CUDA graph
__global__ void adam_lr_kernel(float* state, float lr, float b1, float b2) {
  const float t = state[0] + 1.f;
  state[0] = t;
  const double c1 = 1.0 - pow((double)b1, (double)t), c2 = 1.0 - pow((double)b2, (double)t);
  state[1] = (float)((double)lr * sqrt(c2) / c1);
}
extern "C" int bg_adam_lr_step(float* state, float lr, float beta1, float beta2, void* stream) {
  adam_lr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, lr, beta1, beta2);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
