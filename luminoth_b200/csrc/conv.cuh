// Convolution layer descriptors shared by the engine and the stand-alone op.
#pragma once
#include "common.cuh"
#include <vector>

namespace lumi {

// One conv layer, device-resident, both weight forms.
struct ConvLayer {
  int kh = 1, kw = 1, cin = 0, cout = 0;
  int stride = 1, rate = 1;
  int act = ACT_NONE;
  // SIMT form: TF layout [kh*kw*cin][cout] fp32, epilogue v = acc*scale + bias
  float* w_f32 = nullptr;
  float* scale = nullptr;   // [cout]  (BN gamma*rsqrt(var+eps), or 1)
  float* bias = nullptr;    // [cout]  (BN beta - mean*scale, or conv bias)
  // tcgen05 form: [cout_pad][kh*kw*cin] fp16 hi/lo planes of w * 2^e[c]; scale_tc = scale * 2^-e[c]
  __half* w_hi = nullptr;
  __half* w_lo = nullptr;
  float* scale_tc = nullptr;
  int cout_pad = 0;
  bool tc_ready = false;
};

// Scratch of the stream-K schedule (see conv.cu): one partial-accumulator slot + flag per persistent CTA.
// One workspace per stream: launches that share it must be stream-ordered.
struct ConvWorkspace {
  float* partials = nullptr;
  int* flags = nullptr;
  unsigned epoch = 0;
  int ctas = 0;
};
void conv_workspace_create(ConvWorkspace& w);
void conv_workspace_free(ConvWorkspace& w);
// Per-device facts (SM count, opt-in dynamic shared memory of a kernel) are cached per CUDA device, never per
// process: two engines on two devices, or on two threads, share no mutable launch state.
int device_sm_count();            // SMs of the CURRENT device
constexpr int LUMI_MAX_DEVICES = 64;

struct ConvIO {
  Act in;
  Act out;                  // split-plane output (used when out_f32 == nullptr)
  float* out_f32 = nullptr; // optional fp32 NHWC output [n,ho,wo,cout]
  Act res;                  // optional residual (res.hi == nullptr -> none)
  int res_stride = 1;       // residual sampled at (oy*res_stride, ox*res_stride)  (slim `subsample`)
  int pad_t = 0, pad_l = 0;
  int ho = 0, wo = 0;
  int* overflow_flag = nullptr;
  ConvWorkspace* sk = nullptr;  // enables stream-K scheduling on the tcgen05 path (nullptr: whole tiles only)
  int streamk = 1;              // stream-K policy of THIS launch: 0 off, 1 auto (wave-quantisation heuristic), 2 whenever possible
  int chunk_tail = 2;           // stages per D1 chunk after the first eight stages of a tile (1, 2 or 4; see tc_chunk_end)
  int cta2 = 0;                 // CTA-pair (cta_group::2) kernel on residual-free layers with at least this many K stages per tile (0 = never)
  int halo = 0;                 // halo-patch kernels on 3x3 stride-1 layers: 0 never, 1 single CTA, 2 CTA pairs where C_out % 128 == 0
  int halo_tiles_pct = 150;     // ... while their M-tile count stays within this percentage of the generic kernel's
  int halo_baseoff = 0;         // (bring-up switch) write the patch views' swizzle phase into the matrix descriptors
  int epi16 = 0;                // 16-epilogue-warp kernels on layers with at most this many K stages per tile (0 = never)
  int sm_reserve = 0;           // SMs a persistent launch leaves free (the engine's two-stream pipeline sets 8)
  // Optional strided ("Toeplitz") view of the input for the tcgen05 path: element pitches between
  // consecutive pixels / rows / images (0 = dense NHWC).  Used by the space-to-depth stem, where each
  // A row is the 64 contiguous fp16 of four horizontally adjacent 16-channel pixels.
  long in_pix_pitch = 0, in_row_pitch = 0, in_img_pitch = 0;
};

// host-side packing (w: TF layout on host)
void conv_layer_upload(ConvLayer& L, const float* w_host, const float* scale_host, const float* bias_host);
void conv_layer_free(ConvLayer& L);

bool conv_tc_supported(const ConvLayer& L, const ConvIO& io);
void launch_conv_simt(const ConvLayer& L, const ConvIO& io, cudaStream_t st);
void launch_conv_tc(const ConvLayer& L, const ConvIO& io, cudaStream_t st);

// TF padding arithmetic
inline void tf_same(int in, int k, int stride, int rate, int& out, int& pad_before) {
  int keff = k + (k - 1) * (rate - 1);
  out = (in + stride - 1) / stride;
  int total = (out - 1) * stride + keff - in;
  if (total < 0) total = 0;
  pad_before = total / 2;
}
inline int tf_valid(int in, int k, int stride, int rate) {
  int keff = k + (k - 1) * (rate - 1);
  return (in - keff) / stride + 1;
}

}  // namespace lumi
