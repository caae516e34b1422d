// Gather-GEMM geometry: every conv / transposed conv / dgrad / wgrad of the BigGAN step is described
// as "phases x taps" (see bg_common.cuh).  Reference semantics: ops.py:49-113 (conv: explicit
// tf.pad + VALID conv2d), ops.py:116-139 (deconv: conv2d_transpose SAME).
#include <stdarg.h>
#include "bg_common.cuh"

static thread_local char g_err[512] = "";

void bg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* bg_last_error(void) { return g_err; }
extern "C" int bg_abi_version(void) { return 1; }

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline int pymod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

// transposed-conv style gather: out[o] = sum_a in[(o + off - a)/s], phases over o mod s
static int build_tconv_phases(BgGeom* g, int Hout, int Wout, int kh, int kw, int s, int off_y, int off_x) {
  if (s > 2) { bg_set_error("stride %d > 2 unsupported", s); return BG_ERR_UNSUPPORTED; }
  g->ostep = s;
  g->istep = 1;
  g->nphase = 0;
  for (int py = 0; py < s; ++py)
    for (int px = 0; px < s; ++px) {
      BgPhase& p = g->ph[g->nphase];
      p.oy0 = py; p.ox0 = px;
      p.Hp = (Hout - py + s - 1) / s;
      p.Wp = (Wout - px + s - 1) / s;
      p.ntaps = 0;
      for (int a = 0; a < kh; ++a) {
        if (pymod(py + off_y - a, s) != 0) continue;
        for (int c = 0; c < kw; ++c) {
          if (pymod(px + off_x - c, s) != 0) continue;
          if (p.ntaps >= BG_MAX_TAPS) { bg_set_error("more than %d taps per phase (k=%dx%d s=%d)", BG_MAX_TAPS, kh, kw, s); return BG_ERR_UNSUPPORTED; }
          BgTap& t = p.taps[p.ntaps++];
          t.dy = (short)floordiv(py + off_y - a, s);
          t.dx = (short)floordiv(px + off_x - c, s);
          t.widx = a * kw + c;
        }
      }
      if (p.Hp > 0 && p.Wp > 0) g->nphase++;
    }
  return BG_OK;
}

// plain conv style gather: out[o] = sum_a in[o*s + a - p]
static int build_conv_phase(BgGeom* g, int Hout, int Wout, int kh, int kw, int s, int pt, int pl) {
  if (kh * kw > BG_MAX_TAPS) { bg_set_error("kernel %dx%d has more than %d taps", kh, kw, BG_MAX_TAPS); return BG_ERR_UNSUPPORTED; }
  g->ostep = 1;
  g->istep = s;
  g->nphase = 1;
  BgPhase& p = g->ph[0];
  p.oy0 = 0; p.ox0 = 0; p.Hp = Hout; p.Wp = Wout; p.ntaps = 0;
  for (int a = 0; a < kh; ++a)
    for (int c = 0; c < kw; ++c) {
      BgTap& t = p.taps[p.ntaps++];
      t.dy = (short)(a - pt);
      t.dx = (short)(c - pl);
      t.widx = a * kw + c;
    }
  return BG_OK;
}

int bg_build_geom(const bg_conv_desc* d, int op, BgGeom* g) {
  memset(g, 0, sizeof(*g));
  BG_REQUIRE(d->B > 0 && d->Hi > 0 && d->Wi > 0 && d->Ci > 0 && d->Ho > 0 && d->Wo > 0 && d->Co > 0,
             BG_ERR_SHAPE, "conv desc: non-positive dimension");
  BG_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride >= 1 && d->stride <= 2, BG_ERR_UNSUPPORTED,
             "conv desc: kernel %dx%d stride %d unsupported", d->kh, d->kw, d->stride);
  const int s = d->stride;
  g->B = d->B;
  const long tapsz = (long)d->Ci * d->Co;
  g->wtap = tapsz;
  if (!d->transposed) {
    const int Hpad = d->Hi + d->pad_t + d->pad_b, Wpad = d->Wi + d->pad_l + d->pad_r;
    BG_REQUIRE((Hpad - d->kh) / s + 1 == d->Ho && (Wpad - d->kw) / s + 1 == d->Wo, BG_ERR_SHAPE,
               "conv desc: output %dx%d inconsistent with input %dx%d k%d s%d pads", d->Ho, d->Wo, d->Hi, d->Wi, d->kh, s);
    if (d->pad_mode == BG_PAD_REFLECT)
      BG_REQUIRE(d->pad_t < d->Hi && d->pad_b < d->Hi && d->pad_l < d->Wi && d->pad_r < d->Wi, BG_ERR_SHAPE,
                 "reflect pad must be smaller than the input");
    if (op == BG_OP_FWD || op == BG_OP_WGRAD) {
      g->Hi = d->Hi; g->Wi = d->Wi; g->Ci = d->Ci;
      g->Ho = d->Ho; g->Wo = d->Wo; g->Co = d->Co;
      g->pad_mode = d->pad_mode;
      g->wk = d->Co; g->wn = 1;                       // HWIO: (tap, ci, co)
      return build_conv_phase(g, d->Ho, d->Wo, d->kh, d->kw, s, d->pad_t, d->pad_l);
    }
    // DGRAD: gather over dy, write dx (padded size when reflect-padded)
    const bool padded = d->pad_mode == BG_PAD_REFLECT && (d->pad_t | d->pad_b | d->pad_l | d->pad_r);
    g->Hi = d->Ho; g->Wi = d->Wo; g->Ci = d->Co;
    g->Ho = padded ? Hpad : d->Hi; g->Wo = padded ? Wpad : d->Wi; g->Co = d->Ci;
    g->pad_mode = BG_PAD_ZERO;
    g->wk = 1; g->wn = d->Co;                         // HWIO read as (tap, k=co, n=ci)
    return build_tconv_phases(g, g->Ho, g->Wo, d->kh, d->kw, s, padded ? 0 : d->pad_t, padded ? 0 : d->pad_l);
  }
  // conv2d_transpose SAME
  BG_REQUIRE(d->Ho == d->Hi * s && d->Wo == d->Wi * s, BG_ERR_SHAPE, "deconv desc: Ho must be Hi*stride");
  if (op == BG_OP_FWD) {
    g->Hi = d->Hi; g->Wi = d->Wi; g->Ci = d->Ci;
    g->Ho = d->Ho; g->Wo = d->Wo; g->Co = d->Co;
    g->pad_mode = BG_PAD_ZERO;
    g->wk = 1; g->wn = d->Ci;                         // [kh,kw,Co,Ci] read as (tap, k=ci, n=co)
    return build_tconv_phases(g, d->Ho, d->Wo, d->kh, d->kw, s, d->pad_t, d->pad_l);
  }
  // DGRAD / WGRAD of the transposed conv: plain strided conv over dy
  g->Hi = d->Ho; g->Wi = d->Wo; g->Ci = d->Co;
  g->Ho = d->Hi; g->Wo = d->Wi; g->Co = d->Ci;
  g->pad_mode = BG_PAD_ZERO;
  g->wk = d->Ci; g->wn = 1;                           // [kh,kw,Co,Ci] read as (tap, k=co, n=ci)
  return build_conv_phase(g, d->Hi, d->Wi, d->kh, d->kw, s, d->pad_t, d->pad_l);
}
