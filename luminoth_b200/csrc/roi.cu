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

__global__ void __launch_bounds__(256) roi_pool_kernel(const RoiArgs a) {
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
#pragma unroll
      for (int sy = 0; sy < 2; ++sy) {
        const int cy = py * 2 + sy;
        const float in_y = a.crop_h > 1 ? __fadd_rn(__fmul_rn(y1, Hm1), __fmul_rn((float)cy, hs))
                                        : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(y1, y2)), Hm1);
        const bool y_ok = !(in_y < 0.f || in_y > Hm1);
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const float yl = __fsub_rn(in_y, (float)top);
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
          const int cx = px * 2 + sx;
          const float in_x = a.crop_w > 1 ? __fadd_rn(__fmul_rn(x1, Wm1), __fmul_rn((float)cx, ws))
                                          : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x1, x2)), Wm1);
          const bool ok = y_ok && !(in_x < 0.f || in_x > Wm1);
          if (!ok) {              // extrapolation_value = 0
#pragma unroll
            for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], 0.f);
            continue;
          }
          const int lef = (int)floorf(in_x), rig = (int)ceilf(in_x);
          const float xl = __fsub_rn(in_x, (float)lef);
          float tl[8], tr[8], bl[8], br[8];
          load8(fhi, flo, ((size_t)top * a.fw + lef) * a.c + c0, tl);
          load8(fhi, flo, ((size_t)top * a.fw + rig) * a.c + c0, tr);
          load8(fhi, flo, ((size_t)bot * a.fw + lef) * a.c + c0, bl);
          load8(fhi, flo, ((size_t)bot * a.fw + rig) * a.c + c0, br);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float tv = __fadd_rn(tl[j], __fmul_rn(__fsub_rn(tr[j], tl[j]), xl));
            const float bv = __fadd_rn(bl[j], __fmul_rn(__fsub_rn(br[j], bl[j]), xl));
            best[j] = fmaxf(best[j], __fadd_rn(tv, __fmul_rn(__fsub_rn(bv, tv), yl)));
          }
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
