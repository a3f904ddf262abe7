// ROI crop (bilinear, tf.image.crop_and_resize) fused with the 2x2/2 max pool
// -- luminoth/models/fasterrcnn/roi_pool.py:37-95 (quirks Q3/Q4: boxes are
// normalised by the IMAGE size, crop size is [2*pooled_width, 2*pooled_height]).
//
// One CTA = one ROI x 256-channel slice; 8 warps walk the pooled cells, a lane
// owns 8 channels (one 16 B vector per fp16 plane), so the ROI's footprint of
// the feature map is re-read from L1, not L2/HBM.  HBM-bound on the output write.
#include "ops.cuh"

namespace lumi {

struct RoiArgs {
  const __half* fhi; const __half* flo;
  int n, fh, fw, c;
  const float* rois; const int* counts; int rmax;
  float im_h, im_w;
  int crop_h, crop_w;        // 2*pw, 2*ph  (sic)
  __half* ohi; __half* olo;  // (n*rmax, crop_h/2, crop_w/2, c) or nullptr (mean only)
  __half* mhi; __half* mlo;  // optional fused tf.reduce_mean over the pooled cells: (n*rmax, c)
};

__device__ __forceinline__ void load8(const __half* hi, const __half* lo, size_t off, float (&v)[8]) {
  uint4 vh = __ldg(reinterpret_cast<const uint4*>(hi + off));
  uint4 vl = __ldg(reinterpret_cast<const uint4*>(lo + off));
  const __half* ph = reinterpret_cast<const __half*>(&vh);
  const __half* pl = reinterpret_cast<const __half*>(&vl);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = join_f16(ph[j], pl[j]);
}

// horizontal lerp of one feature row at the two x samples of a pooled cell; column loads are shared
// between the two samples whenever they hit the same feature cell (all indices are warp-uniform).
__device__ __forceinline__ void row_interp(const __half* fhi, const __half* flo, size_t rowbase, int c, int c0,
                                           const int (&lef)[2], const int (&rig)[2], const float (&xl)[2],
                                           float (&h0)[8], float (&h1)[8]) {
  float l0[8], r0[8], l1[8], r1[8];
  load8(fhi, flo, (rowbase + lef[0]) * c + c0, l0);
  if (rig[0] != lef[0]) load8(fhi, flo, (rowbase + rig[0]) * c + c0, r0);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) r0[j] = l0[j];
  }
  if (lef[1] == lef[0]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) l1[j] = l0[j];
  } else if (lef[1] == rig[0]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) l1[j] = r0[j];
  } else {
    load8(fhi, flo, (rowbase + lef[1]) * c + c0, l1);
  }
  if (rig[1] == rig[0]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r1[j] = r0[j];
  } else if (rig[1] == lef[1]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r1[j] = l1[j];
  } else {
    load8(fhi, flo, (rowbase + rig[1]) * c + c0, r1);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h0[j] = __fadd_rn(l0[j], __fmul_rn(__fsub_rn(r0[j], l0[j]), xl[0]));
    h1[j] = __fadd_rn(l1[j], __fmul_rn(__fsub_rn(r1[j], l1[j]), xl[1]));
  }
}

__global__ void __launch_bounds__(256, 2) roi_pool_kernel(const RoiArgs a) {
  const int row = blockIdx.x;                 // global roi row = img*rmax + r
  const int img = row / a.rmax, r = row % a.rmax;
  const int cslice = blockIdx.y * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = cslice + lane * 8;
  const int oh = a.crop_h >> 1, ow = a.crop_w >> 1;
  const bool live = (a.counts == nullptr || r < a.counts[img]) && c0 < a.c;
  const size_t obase = (size_t)row * oh * ow * a.c;
  // normalised box, TF order (y1,x1,y2,x2)
  const float* rb = a.rois + (size_t)row * 4;
  const float x1 = __fdiv_rn(rb[0], a.im_w), y1 = __fdiv_rn(rb[1], a.im_h);
  const float x2 = __fdiv_rn(rb[2], a.im_w), y2 = __fdiv_rn(rb[3], a.im_h);
  const float Hm1 = (float)(a.fh - 1), Wm1 = (float)(a.fw - 1);
  const float hs = a.crop_h > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(y2, y1), Hm1), (float)(a.crop_h - 1)) : 0.f;
  const float ws = a.crop_w > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(x2, x1), Wm1), (float)(a.crop_w - 1)) : 0.f;
  const __half* fhi = a.fhi + (size_t)img * a.fh * a.fw * a.c;
  const __half* flo = a.flo + (size_t)img * a.fh * a.fw * a.c;

  float msum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) msum[j] = 0.f;
  for (int cell = warp; cell < oh * ow; cell += 8) {
    if (c0 >= a.c) break;
    const int py = cell / ow, px = cell % ow;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = live ? -INFINITY : 0.f;
    if (live) {
      // sample coordinates of the 2x2 crop samples under this pooled cell (TF crop_and_resize arithmetic)
      int top[2], bot[2], lef[2], rig[2];
      float yl[2], xl[2];
      bool y_ok[2], x_ok[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float in_y = a.crop_h > 1 ? __fadd_rn(__fmul_rn(y1, Hm1), __fmul_rn((float)(py * 2 + s), hs))
                                        : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(y1, y2)), Hm1);
        y_ok[s] = !(in_y < 0.f || in_y > Hm1);
        top[s] = y_ok[s] ? (int)floorf(in_y) : 0;
        bot[s] = y_ok[s] ? (int)ceilf(in_y) : 0;
        yl[s] = __fsub_rn(in_y, (float)top[s]);
        const float in_x = a.crop_w > 1 ? __fadd_rn(__fmul_rn(x1, Wm1), __fmul_rn((float)(px * 2 + s), ws))
                                        : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x1, x2)), Wm1);
        x_ok[s] = !(in_x < 0.f || in_x > Wm1);
        lef[s] = x_ok[s] ? (int)floorf(in_x) : 0;
        rig[s] = x_ok[s] ? (int)ceilf(in_x) : 0;
        xl[s] = __fsub_rn(in_x, (float)lef[s]);
      }
      // row-interpolated values, shared between the two y samples when they touch the same feature rows
      float ht[2][8], hb[2][8];          // [sx][ch] for the current sy: top row / bottom row
      float pt[2][8], pb[2][8];          // previous sy (sy = 0)
#pragma unroll
      for (int sy = 0; sy < 2; ++sy) {
        if (sy == 0) {
          row_interp(fhi, flo, (size_t)top[0] * a.fw, a.c, c0, lef, rig, xl, ht[0], ht[1]);
          if (bot[0] != top[0]) row_interp(fhi, flo, (size_t)bot[0] * a.fw, a.c, c0, lef, rig, xl, hb[0], hb[1]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { hb[0][j] = ht[0][j]; hb[1][j] = ht[1][j]; }
          }
        } else {
          if (top[1] == top[0]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ht[0][j] = pt[0][j]; ht[1][j] = pt[1][j]; }
          } else if (top[1] == bot[0]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ht[0][j] = pb[0][j]; ht[1][j] = pb[1][j]; }
          } else {
            row_interp(fhi, flo, (size_t)top[1] * a.fw, a.c, c0, lef, rig, xl, ht[0], ht[1]);
          }
          if (bot[1] == bot[0]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { hb[0][j] = pb[0][j]; hb[1][j] = pb[1][j]; }
          } else if (bot[1] == top[1]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { hb[0][j] = ht[0][j]; hb[1][j] = ht[1][j]; }
          } else {
            row_interp(fhi, flo, (size_t)bot[1] * a.fw, a.c, c0, lef, rig, xl, hb[0], hb[1]);
          }
        }
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
          const bool ok = y_ok[sy] && x_ok[sx];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = ok ? __fadd_rn(ht[sx][j], __fmul_rn(__fsub_rn(hb[sx][j], ht[sx][j]), yl[sy])) : 0.f;
            best[j] = fmaxf(best[j], v);              // extrapolation_value = 0
          }
        }
        if (sy == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { pt[0][j] = ht[0][j]; pt[1][j] = ht[1][j]; pb[0][j] = hb[0][j]; pb[1][j] = hb[1][j]; }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) msum[j] += best[j];
    if (a.ohi) {
      uint4 vh, vl;
      __half* qh = reinterpret_cast<__half*>(&vh);
      __half* ql = reinterpret_cast<__half*>(&vl);
#pragma unroll
      for (int j = 0; j < 8; ++j) split_f32(best[j], qh[j], ql[j]);
      const size_t off = obase + (size_t)cell * a.c + c0;
      *reinterpret_cast<uint4*>(a.ohi + off) = vh;
      *reinterpret_cast<uint4*>(a.olo + off) = vl;
    }
  }
  if (a.mhi) {            // fused spatial mean (rcnn.py:188): warp partials -> fixed-order sum -> / cells
    __shared__ float part[8][256];
#pragma unroll
    for (int j = 0; j < 8; ++j) part[warp][lane * 8 + j] = msum[j];
    __syncthreads();
    const int ch = cslice + threadIdx.x;
    if (ch < a.c) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += part[w8][threadIdx.x];
      const float v = __fdiv_rn(s, (float)(oh * ow));
      __half h, l;
      split_f32(v, h, l);
      a.mhi[(size_t)row * a.c + ch] = h;
      a.mlo[(size_t)row * a.c + ch] = l;
    }
  }
}

void launch_roi_pool(Act fmap, const float* rois, const int* counts, int rmax, float im_h, float im_w, int ph, int pw,
                     Act out, Act mean, cudaStream_t st) {
  LUMI_REQUIRE(fmap.c % 8 == 0, "roi_pool: C must be a multiple of 8");
  RoiArgs a;
  a.fhi = fmap.hi; a.flo = fmap.lo; a.n = fmap.n; a.fh = fmap.h; a.fw = fmap.w; a.c = fmap.c;
  a.rois = rois; a.counts = counts; a.rmax = rmax; a.im_h = im_h; a.im_w = im_w;
  a.crop_h = pw * 2; a.crop_w = ph * 2;      // roi_pool.py:77 passes [pooled_width*2, pooled_height*2]
  a.ohi = out.hi; a.olo = out.lo;
  a.mhi = mean.hi; a.mlo = mean.lo;
  LUMI_REQUIRE(out.hi || mean.hi, "roi_pool: no output requested");
  long rows = (long)fmap.n * rmax;
  if (!rows) return;
  dim3 grid((unsigned)rows, (unsigned)cdiv(fmap.c, 256));
  roi_pool_kernel<<<grid, 256, 0, st>>>(a);
  count_launch();
  LUMI_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lumi
