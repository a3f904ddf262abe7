// ROI crop (bilinear, tf.image.crop_and_resize) fused with the 2x2/2 max pool
// -- luminoth/models/fasterrcnn/roi_pool.py:37-95 (quirks Q3/Q4: boxes are
// normalised by the IMAGE size, crop size is [2*pooled_width, 2*pooled_height]).
//
// One CTA = one ROI x 256-channel slice; 8 warps walk the pooled cells, a lane owns 8 channels.
// The gather reads an fp32 copy of the feature map (made once per forward) so a tap costs two 16 B
// loads and no conversions; taps shared by the 2x2 samples of a cell are loaded once.  The kernel is
// instruction / L1-bound (12.8 G bilinear taps per batch at R = 2000), not HBM-bound.
#include "ops.cuh"
#include <cstdlib>

namespace lumi {

struct RoiArgs {
  const float* fmap;         // fp32 NHWC copy of the feature map (one-off conversion; saves 3 instr / element / tap)
  int n, fh, fw, c;
  const float* rois; const int* counts; int rmax;
  float im_h, im_w;
  int crop_h, crop_w;        // 2*pw, 2*ph  (sic)
  __half* ohi; __half* olo;  // (n*rmax, crop_h/2, crop_w/2, c) or nullptr (mean only)
  __half* mhi; __half* mlo;  // optional fused tf.reduce_mean over the pooled cells: (n*rmax, c)
};

struct Samp { int lo, hi; float lerp; int ok; };    // one crop sample coordinate along y or x

template <int CPL>
__device__ __forceinline__ void load8f(const float* f, size_t off, float (&v)[CPL]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(f + off));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  if (CPL == 8) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(f + off) + 1);
    v[CPL - 4] = b.x; v[CPL - 3] = b.y; v[CPL - 2] = b.z; v[CPL - 1] = b.w;
  }
}

// horizontal lerp of one feature row at the two x samples of a pooled cell; column loads are shared
// between the two samples whenever they hit the same feature cell (all indices are warp-uniform).
template <int CPL>
__device__ __forceinline__ void lerp8(const float (&l)[CPL], const float (&r)[CPL], float t, float (&h)[CPL]) {
#pragma unroll
  for (int j = 0; j < CPL; ++j) h[j] = fmaf(r[j] - l[j], t, l[j]);     // left + (right - left) * lerp
}
// Every re-use case names its source registers statically, so sharing a column costs no register moves (copying the
// shared values into l1 / r1 first measured 2.48 ms vs 2.38 ms per step).  The same treatment of the row-level sharing
// in the caller is register-neutral on its own (2.40 ms) but spills when combined with this one (3.43 ms): not done.
template <int CPL>
__device__ __forceinline__ void row_interp(const float* f, size_t rowbase, int c, int c0, const Samp& s0,
                                           const Samp& s1, float (&h0)[CPL], float (&h1)[CPL]) {
  float l0[CPL], r0[CPL], l1[CPL], r1[CPL];
  const float* row = f + rowbase * c + c0;
  load8f<CPL>(row, (size_t)s0.lo * c, l0);
  const bool r0_is_l0 = s0.hi == s0.lo;
  if (!r0_is_l0) { load8f<CPL>(row, (size_t)s0.hi * c, r0); lerp8<CPL>(l0, r0, s0.lerp, h0); }
  else lerp8<CPL>(l0, l0, s0.lerp, h0);
  if (s1.lo == s0.lo) {                                        // left1 = left0
    if (s1.hi == s0.hi) { if (r0_is_l0) lerp8<CPL>(l0, l0, s1.lerp, h1); else lerp8<CPL>(l0, r0, s1.lerp, h1); }
    else if (s1.hi == s1.lo) lerp8<CPL>(l0, l0, s1.lerp, h1);
    else { load8f<CPL>(row, (size_t)s1.hi * c, r1); lerp8<CPL>(l0, r1, s1.lerp, h1); }
  } else if (s1.lo == s0.hi) {                                 // left1 = right0 (r0 is loaded: s0.hi != s0.lo here)
    if (s1.hi == s0.hi || s1.hi == s1.lo) lerp8<CPL>(r0, r0, s1.lerp, h1);
    else { load8f<CPL>(row, (size_t)s1.hi * c, r1); lerp8<CPL>(r0, r1, s1.lerp, h1); }
  } else {
    load8f<CPL>(row, (size_t)s1.lo * c, l1);
    if (s1.hi == s0.hi) { if (r0_is_l0) lerp8<CPL>(l1, l0, s1.lerp, h1); else lerp8<CPL>(l1, r0, s1.lerp, h1); }
    else if (s1.hi == s1.lo) lerp8<CPL>(l1, l1, s1.lerp, h1);
    else { load8f<CPL>(row, (size_t)s1.hi * c, r1); lerp8<CPL>(l1, r1, s1.lerp, h1); }
  }
}

// RB consecutive ROIs per CTA: their RB*49 pooled cells are dealt round-robin to the 8 warps.  With one ROI per
// CTA, 49 cells on 8 warps leave seven warps idle for 1/8 of the CTA's life (ncu: 10 % of all samples stalled at the
// final barrier); with four ROIs the imbalance is 196 = 8*24 + 4 -> 2 %.
template <int CPL, int RB, int NW>
__global__ void __launch_bounds__(32 * NW, (CPL == 8 ? 2 : 3) * (8 / NW)) roi_pool_kernel(const RoiArgs a) {
  constexpr int SLICE = 32 * CPL;             // channels per CTA (one warp-wide vector of CPL channels per lane)
  const int row0 = blockIdx.x * RB;           // first global roi row (= img*rmax + r) of this CTA
  const int rows_total = a.n * a.rmax;
  const int cslice = blockIdx.y * SLICE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = cslice + lane * CPL;
  const int oh = a.crop_h >> 1, ow = a.crop_w >> 1;
  const int ncell = oh * ow;

  // sample tables: crop_h y-samples then crop_w x-samples (TF crop_and_resize arithmetic, once per CTA and ROI)
  __shared__ Samp samp_all[RB][64];
  const int nsamp = a.crop_h + a.crop_w;
  if ((int)threadIdx.x < RB * nsamp && row0 + (int)threadIdx.x / nsamp < rows_total) {
    const int rl_ = threadIdx.x / nsamp, ts = threadIdx.x % nsamp;
    const int row = row0 + rl_;
    const bool is_y = ts < a.crop_h;
    const int k = is_y ? ts : ts - a.crop_h;
    const float* rb = a.rois + (size_t)row * 4;
    // normalised box, TF order (y1,x1,y2,x2): divided by the IMAGE size (quirk Q3)
    const float lo_n = is_y ? __fdiv_rn(rb[1], a.im_h) : __fdiv_rn(rb[0], a.im_w);
    const float hi_n = is_y ? __fdiv_rn(rb[3], a.im_h) : __fdiv_rn(rb[2], a.im_w);
    const int crop = is_y ? a.crop_h : a.crop_w;
    const float Dm1 = (float)((is_y ? a.fh : a.fw) - 1);
    const float step = crop > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(hi_n, lo_n), Dm1), (float)(crop - 1)) : 0.f;
    const float in = crop > 1 ? __fadd_rn(__fmul_rn(lo_n, Dm1), __fmul_rn((float)k, step))
                              : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(lo_n, hi_n)), Dm1);
    Samp s;
    s.ok = !(in < 0.f || in > Dm1);
    s.lo = s.ok ? (int)floorf(in) : 0;
    s.hi = s.ok ? (int)ceilf(in) : 0;
    s.lerp = __fsub_rn(in, (float)s.lo);
    samp_all[rl_][ts] = s;
  }
  __syncthreads();

  __shared__ float part[RB][NW][SLICE];       // per-warp partial sums of the fused spatial mean
  float msum[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) msum[j] = 0.f;
  for (int rl = 0; rl < RB; ++rl) {
    const int row = row0 + rl;
    if (row >= rows_total) break;
    const int img = row / a.rmax, r = row - img * a.rmax;
    const bool live = (a.counts == nullptr || r < a.counts[img]) && c0 < a.c;
    const size_t obase = (size_t)row * ncell * a.c;
    const float* f = a.fmap + (size_t)img * a.fh * a.fw * a.c;
    const Samp* samp = samp_all[rl];
    // the RB*ncell cells of the CTA are dealt round-robin: this warp's first cell inside ROI rl
    const int first = (((warp - rl * ncell) % NW) + NW) % NW;
  for (int cell = first; cell < ncell; cell += NW) {
    if (c0 >= a.c) break;
    const int py = cell / ow, px = cell % ow;
    float best[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) best[j] = live ? -INFINITY : 0.f;
    if (live) {
      const Samp y0 = samp[py * 2], y1 = samp[py * 2 + 1];
      const Samp x0 = samp[a.crop_h + px * 2], x1 = samp[a.crop_h + px * 2 + 1];
      // row-interpolated values, shared between the two y samples when they touch the same feature rows
      float t0[2][CPL], b0[2][CPL];          // sample row 0: top / bottom feature row, [sx][ch]
      row_interp<CPL>(f, (size_t)y0.lo * a.fw, a.c, c0, x0, x1, t0[0], t0[1]);
      if (y0.hi != y0.lo) row_interp<CPL>(f, (size_t)y0.hi * a.fw, a.c, c0, x0, x1, b0[0], b0[1]);
      else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { b0[0][j] = t0[0][j]; b0[1][j] = t0[1][j]; }
      }
#pragma unroll
      for (int sx = 0; sx < 2; ++sx) {
        const bool ok = y0.ok && (sx ? x1.ok : x0.ok);          // warp-uniform
        if (ok) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) best[j] = fmaxf(best[j], fmaf(b0[sx][j] - t0[sx][j], y0.lerp, t0[sx][j]));
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) best[j] = fmaxf(best[j], 0.f);              // extrapolation_value = 0
        }
      }
      float t1[2][CPL], b1[2][CPL];
      if (y1.lo == y0.lo) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { t1[0][j] = t0[0][j]; t1[1][j] = t0[1][j]; }
      } else if (y1.lo == y0.hi) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { t1[0][j] = b0[0][j]; t1[1][j] = b0[1][j]; }
      } else {
        row_interp<CPL>(f, (size_t)y1.lo * a.fw, a.c, c0, x0, x1, t1[0], t1[1]);
      }
      if (y1.hi == y0.hi) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { b1[0][j] = b0[0][j]; b1[1][j] = b0[1][j]; }
      } else if (y1.hi == y1.lo) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) { b1[0][j] = t1[0][j]; b1[1][j] = t1[1][j]; }
      } else {
        row_interp<CPL>(f, (size_t)y1.hi * a.fw, a.c, c0, x0, x1, b1[0], b1[1]);
      }
#pragma unroll
      for (int sx = 0; sx < 2; ++sx) {
        const bool ok = y1.ok && (sx ? x1.ok : x0.ok);
        if (ok) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) best[j] = fmaxf(best[j], fmaf(b1[sx][j] - t1[sx][j], y1.lerp, t1[sx][j]));
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) best[j] = fmaxf(best[j], 0.f);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) msum[j] += best[j];
    if (a.ohi) {
      uint4 vh, vl;
      __half* qh = reinterpret_cast<__half*>(&vh);
      __half* ql = reinterpret_cast<__half*>(&vl);
#pragma unroll
      for (int j = 0; j < CPL; ++j) split_f32(best[j], qh[j], ql[j]);
      const size_t off = obase + (size_t)cell * a.c + c0;
      if (CPL == 8) {
        *reinterpret_cast<uint4*>(a.ohi + off) = vh;
        *reinterpret_cast<uint4*>(a.olo + off) = vl;
      } else {
        *reinterpret_cast<uint2*>(a.ohi + off) = make_uint2(vh.x, vh.y);
        *reinterpret_cast<uint2*>(a.olo + off) = make_uint2(vl.x, vl.y);
      }
    }
  }
    if (a.mhi) {          // this warp's share of ROI rl
#pragma unroll
      for (int j = 0; j < CPL; ++j) { part[rl][warp][lane * CPL + j] = msum[j]; msum[j] = 0.f; }
    }
  }
  if (a.mhi) {            // fused spatial mean (rcnn.py:188): warp partials -> fixed-order sum -> / cells
    __syncthreads();
    for (int i = threadIdx.x; i < SLICE; i += 32 * NW) {
      const int ch = cslice + i;
      if (ch >= a.c) break;
      for (int rl = 0; rl < RB && row0 + rl < rows_total; ++rl) {
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) s += part[rl][w8][i];
        const float v = __fdiv_rn(s, (float)ncell);
        __half h, l;
        split_f32(v, h, l);
        a.mhi[(size_t)(row0 + rl) * a.c + ch] = h;
        a.mlo[(size_t)(row0 + rl) * a.c + ch] = l;
      }
    }
  }
}

void launch_roi_pool(const float* fmap_f32, int n, int fh, int fw, int c, const float* rois, const int* counts, int rmax,
                     float im_h, float im_w, int ph, int pw, Act out, Act mean, cudaStream_t st) {
  LUMI_REQUIRE(c % 8 == 0, "roi_pool: C must be a multiple of 8");
  LUMI_REQUIRE(2 * (ph + pw) <= 64, "roi_pool: pooled size too large");
  RoiArgs a;
  a.fmap = fmap_f32; a.n = n; a.fh = fh; a.fw = fw; a.c = c;
  a.rois = rois; a.counts = counts; a.rmax = rmax; a.im_h = im_h; a.im_w = im_w;
  a.crop_h = pw * 2; a.crop_w = ph * 2;      // roi_pool.py:77 passes [pooled_width*2, pooled_height*2]
  a.ohi = out.hi; a.olo = out.lo;
  a.mhi = mean.hi; a.mlo = mean.lo;
  LUMI_REQUIRE(out.hi || mean.hi, "roi_pool: no output requested");
  long rows = (long)n * rmax;
  if (!rows) return;
  static const int cpl = [] { const char* e = getenv("LUMI_ROI_CPL"); return (e && atoi(e) == 4) ? 4 : 8; }();
  // (measured alternatives at R = 2000, batch 8: 4 channels/lane 3.1 ms, straight-line 16-tap loads
  //  without sharing 3.0 ms, this kernel 2.7 ms)
  static const int rb = [] { const char* e = getenv("LUMI_ROI_RB"); return (e && atoi(e) == 1) ? 1 : 4; }();
  LUMI_REQUIRE(pw * 2 + ph * 2 <= 64, "roi_pool: pooled size too large");
  // 4 warps x 4 ROIs: 196 cells = 49 per warp, no tail at all (measured 2.49 vs 2.51 ms with 8 warps, 2.73 ms with
  // one ROI per CTA)
  static const int nw = [] { const char* e = getenv("LUMI_ROI_NW"); return (e && atoi(e) == 8) ? 8 : 4; }();
  if (cpl == 8) {
    if (rb == 4 && nw == 4 && 4 * (a.crop_h + a.crop_w) <= 128) {
      dim3 grid((unsigned)cdiv64(rows, 4), (unsigned)cdiv(c, 256));
      roi_pool_kernel<8, 4, 4><<<grid, 128, 0, st>>>(a);
    } else if (rb == 4 && 4 * (a.crop_h + a.crop_w) <= 256) {
      dim3 grid((unsigned)cdiv64(rows, 4), (unsigned)cdiv(c, 256));
      roi_pool_kernel<8, 4, 8><<<grid, 256, 0, st>>>(a);
    } else {
      dim3 grid((unsigned)rows, (unsigned)cdiv(c, 256));
      roi_pool_kernel<8, 1, 8><<<grid, 256, 0, st>>>(a);
    }
  } else {                 // 4 channels per lane: half the registers (measured slower: 3.1 vs 2.7 ms at R = 2000)
    dim3 grid((unsigned)rows, (unsigned)cdiv(c, 128));
    roi_pool_kernel<4, 1, 8><<<grid, 256, 0, st>>>(a);
  }
  count_launch();
  LUMI_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lumi
