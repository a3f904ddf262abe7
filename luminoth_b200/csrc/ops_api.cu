// Stand-alone C-ABI operators (per-kernel parity tests and micro-benchmarks).
// Each wraps exactly the launcher the engine uses; temporaries are allocated per
// call and the stream is synchronised before they are released.
#include "../../include/luminoth_b200.h"
#include "conv.cuh"
#include "ops.cuh"

#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

using namespace lumi;

namespace {
thread_local std::string g_op_error;

struct DevBuf {
  void* p = nullptr;
  explicit DevBuf(size_t bytes) { if (bytes) LUMI_CUDA_CHECK(cudaMalloc(&p, bytes)); }
  ~DevBuf() { cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  template <typename T> T* as() { return static_cast<T*>(p); }
};

struct ActBuf {
  DevBuf hi, lo;
  Act a;
  ActBuf(int n, int h, int w, int c) : hi((size_t)n * h * w * c * 2), lo((size_t)n * h * w * c * 2) {
    a.n = n; a.h = h; a.w = w; a.c = c; a.hi = hi.as<__half>(); a.lo = lo.as<__half>();
  }
};

int op_fail(const Error& e) { g_op_error = e.what(); return e.code; }
}  // namespace

#define OP_BEGIN try {
#define OP_END                                           \
  }                                                      \
  catch (const Error& err) { return op_fail(err); }      \
  catch (const std::exception& ex) { g_op_error = ex.what(); return LUMI_EINVAL; }

extern "C" {

const char* lumi_op_last_error(void) { return g_op_error.c_str(); }

int lumi_op_conv2d(const float* x, int n, int h, int w, int cin, const float* wgt, int kh, int kw, int cout, int stride,
                   int rate, int padding, const float* scale, const float* bias, const float* residual, int act,
                   int impl, float* y, int* ho_out, int* wo_out, void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvLayer L;
  L.kh = kh; L.kw = kw; L.cin = cin; L.cout = cout; L.stride = stride; L.rate = rate; L.act = act;
  const size_t nw = (size_t)kh * kw * cin * cout;
  std::vector<float> hw(nw), hs, hb;
  LUMI_CUDA_CHECK(cudaMemcpy(hw.data(), wgt, nw * sizeof(float), cudaMemcpyDeviceToHost));
  if (scale) { hs.resize(cout); LUMI_CUDA_CHECK(cudaMemcpy(hs.data(), scale, cout * sizeof(float), cudaMemcpyDeviceToHost)); }
  if (bias) { hb.resize(cout); LUMI_CUDA_CHECK(cudaMemcpy(hb.data(), bias, cout * sizeof(float), cudaMemcpyDeviceToHost)); }
  conv_layer_upload(L, hw.data(), scale ? hs.data() : nullptr, bias ? hb.data() : nullptr);
  struct Guard { ConvLayer& l; ~Guard() { conv_layer_free(l); } } guard{L};
  int ho, wo, pt = 0, pl = 0;
  if (padding == 1 || (padding == 2 && stride == 1)) {
    tf_same(h, kh, stride, rate, ho, pt); tf_same(w, kw, stride, rate, wo, pl);
  } else if (padding == 2) {
    const int keff = kh + (kh - 1) * (rate - 1);
    pt = pl = (keff - 1) / 2;
    ho = (h + (keff - 1) - keff) / stride + 1; wo = (w + (keff - 1) - keff) / stride + 1;
  } else {
    ho = tf_valid(h, kh, stride, rate); wo = tf_valid(w, kw, stride, rate);
  }
  LUMI_REQUIRE(ho > 0 && wo > 0, "conv2d: empty output");
  if (ho_out) *ho_out = ho;
  if (wo_out) *wo_out = wo;
  if (!y) return LUMI_OK;                 // shape query
  ActBuf in(n, h, w, cin);
  launch_f32_to_act(x, in.a, st);
  ConvIO io;
  io.in = in.a; io.pad_t = pt; io.pad_l = pl; io.ho = ho; io.wo = wo; io.out_f32 = y;
  std::unique_ptr<ActBuf> res;
  if (residual) {
    res.reset(new ActBuf(n, ho, wo, cout));
    launch_f32_to_act(residual, res->a, st);
    io.res = res->a; io.res_stride = 1;
  }
  if (impl >= 1 && impl <= 11) {
    // 1 whole tiles, 2 stream-K forced (fp32 outputs written by the epilogue);
    // 3 / 4 / 5: the engine's inter-layer form -- fp16x2 split planes through the staged TMA-store epilogue (and the
    // TMA-prefetched residual) -- with the 8-warp epilogue (3), the 16-warp short-K kernels allowed (4), and 4 + stream-K (5)
    std::unique_ptr<ActBuf> split_out;
    if (impl >= 3) {
      LUMI_REQUIRE(cout % 32 == 0, "conv2d: split outputs need cout % 32 == 0");
      split_out.reset(new ActBuf(n, ho, wo, cout));
      io.out = split_out->a;
      io.out_f32 = nullptr;
      io.epi16 = (impl == 4 || impl == 5) ? 8 : 0;
      io.cta2 = (impl == 6 || impl == 7) ? 1 : 0;   // 6 / 7: CTA-pair kernel wherever it applies (7: + stream-K forced)
      io.halo = (impl == 8 || impl == 9) ? 1 : ((impl == 10 || impl == 11) ? 2 : 0);   // 8-11: halo-patch kernels (9, 11: + stream-K)
      io.halo_tiles_pct = 1000000;                  // test hook: whenever the shape allows
      if (const char* e = std::getenv("LUMI_HALO_BASEOFF")) io.halo_baseoff = std::atoi(e);
    }
    LUMI_REQUIRE(conv_tc_supported(L, io), "conv2d: this layer shape is not handled by the tensor-core kernel");
    ConvWorkspace sk;
    struct SkGuard { ConvWorkspace& w; ~SkGuard() { conv_workspace_free(w); } } skg{sk};
    if (impl == 2 || impl == 5 || impl == 7 || impl == 9 || impl == 11) {
      conv_workspace_create(sk);
      io.sk = &sk;
      io.streamk = 2;
    }
    launch_conv_tc(L, io, st);
    if (split_out) launch_act_to_f32(split_out->a, y, st);
    LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  } else {
    launch_conv_simt(L, io, st);
  }
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_resize_bilinear(const void* src, int src_is_f32, int h0, int w0, float* dst, int h, int w, void* stream) {
  OP_BEGIN
  LUMI_REQUIRE(src && dst && h0 > 0 && w0 > 0 && h > 0 && w > 0, "resize_bilinear: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_resize_bilinear(src, src_is_f32 != 0, h0, w0, dst, h, w, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_max_pool(const float* x, int n, int h, int w, int c, int k, int stride, int padding, float* y, void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int ho, wo, pt = 0, pl = 0;
  if (padding == 1) { tf_same(h, k, stride, 1, ho, pt); tf_same(w, k, stride, 1, wo, pl); }
  else { ho = tf_valid(h, k, stride, 1); wo = tf_valid(w, k, stride, 1); }
  ActBuf in(n, h, w, c), out(n, ho, wo, c);
  launch_f32_to_act(x, in.a, st);
  launch_max_pool(in.a, out.a, k, stride, pt, pl, st);
  launch_act_to_f32(out.a, y, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_roi_pool(const float* fmap, int n, int fh, int fw, int c, const float* rois, const int32_t* roi_batch, int r,
                     float im_h, float im_w, int ph, int pw, float* y, void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LUMI_REQUIRE(n == 1 || roi_batch == nullptr, "roi_pool op: single image (roi_batch must be NULL or n == 1)");
  (void)roi_batch;
  ActBuf out(r, pw, ph, c);
  launch_roi_pool(fmap, n, fh, fw, c, rois, nullptr, r, im_h, im_w, ph, pw, out.a, Act(), st);
  launch_act_to_f32(out.a, y, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_sort_desc(const float* scores, int n, int32_t* idx_out, void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return LUMI_OK;
  NmsWorkspace ws;
  struct G { NmsWorkspace& w; ~G() { nms_workspace_free(w); } } g{ws};
  nms_workspace_alloc(ws, 1, n, 1);
  launch_sort_desc(scores, n, idx_out, ws, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_nms_sorted(const float* boxes_sorted, int n, float iou_threshold, int max_out, int32_t* keep,
                       int32_t* num_keep, void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LUMI_REQUIRE(n > 0 && max_out > 0, "nms_sorted: n and max_out must be positive");
  NmsWorkspace ws;
  struct G { NmsWorkspace& w; ~G() { nms_workspace_free(w); } } g{ws};
  nms_workspace_alloc(ws, 1, n, max_out);
  launch_nms_sorted(boxes_sorted, n, iou_threshold, max_out, ws, keep, num_keep, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_rpn_proposals(const float* cls_prob, const float* bbox_pred, const float* anchors, int na, float im_h,
                          float im_w, int pre_nms_top_n, int post_nms_top_n, float nms_threshold, float min_prob,
                          int filter_outside, int clip_after_nms, float* proposals, float* scores, int32_t* count,
                          void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LUMI_REQUIRE(na > 0 && pre_nms_top_n > 0 && post_nms_top_n > 0, "rpn_proposals: sizes must be positive");
  NmsWorkspace ws;
  struct G { NmsWorkspace& w; ~G() { nms_workspace_free(w); } } g{ws};
  nms_workspace_alloc(ws, 1, na, post_nms_top_n, pre_nms_top_n < na ? pre_nms_top_n : na);
  RpnParams p{};
  p.na = na; p.im_h = im_h; p.im_w = im_w; p.pre_nms_top_n = pre_nms_top_n; p.post_nms_top_n = post_nms_top_n;
  p.nms_threshold = nms_threshold; p.min_prob = min_prob; p.filter_outside = filter_outside;
  p.clip_after_nms = clip_after_nms; p.apply_nms = 1; p.logits = 0;
  p.cls_stride = 2; p.cls_off = 0; p.box_stride = 4; p.box_off = 0;
  launch_rpn_proposals(cls_prob, bbox_pred, 0, 0, 1, anchors, 1, p, ws, proposals, scores, count, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

int lumi_op_class_detections(const float* boxes_in, const float* deltas, const float* cls_prob, int r, int nc, float im_h,
                             float im_w, float var0, float var1, float min_prob, float nms_threshold, int class_max,
                             int total_max, int ssd_order, float* objects, int32_t* labels, float* probs, int32_t* count,
                             void* stream) {
  OP_BEGIN
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LUMI_REQUIRE(r > 0 && nc > 0 && class_max > 0 && total_max > 0, "class_detections: sizes must be positive");
  NmsWorkspace ws;
  struct G { NmsWorkspace& w; ~G() { nms_workspace_free(w); } } g{ws};
  nms_workspace_alloc(ws, nc, r, class_max);
  DevBuf fk(det_final_scratch_bytes(1, nc, class_max));
  DetParams p{};
  p.r = r; p.nc = nc; p.im_h = im_h; p.im_w = im_w; p.var0 = var0; p.var1 = var1; p.min_prob = min_prob;
  p.nms_threshold = nms_threshold; p.class_max = class_max; p.total_max = total_max;
  p.shared_deltas = ssd_order ? 1 : 0;
  p.prob_stride = nc + 1; p.delta_stride = ssd_order ? 4 : 4 * nc;
  launch_class_detections(boxes_in, (long)r * 4, nullptr, deltas, cls_prob, 1, p, ws, fk.as<float>(), objects, labels,
                          probs, count, st);
  LUMI_CUDA_CHECK(cudaStreamSynchronize(st));
  return LUMI_OK;
  OP_END
}

}  // extern "C"
