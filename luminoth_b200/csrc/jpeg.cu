// JPEG decode on the GPU (the input path of `lumi predict`, predict.py:69-97 opens files with PIL).
// nvJPEG is a CUDA toolkit library (like cuBLAS: a plain library call, not a hand-written kernel); it is bound at
// RUN time with dlopen so that libluminoth_b200.so itself has no link dependency on it -- a box without libnvjpeg
// can still run the engine, and lumi_decode_jpeg reports LUMI_ECUDA there.
#include "../../include/luminoth_b200.h"
#include "common.cuh"

#include <dlfcn.h>
#include <mutex>
#include <nvjpeg.h>

namespace {
thread_local std::string g_jpeg_error;

struct NvJpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*,
                                 int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                           nvjpegImage_t*, cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  std::mutex mu;          // one decode at a time per process (the state object is not re-entrant)
  bool ok = false;
  std::string why;
};

NvJpegApi& api() {
  static NvJpegApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) { a.why = "libnvjpeg.so.12 not found"; return; }
#define LUMI_SYM(field, sym)                                               \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));        \
  if (!a.field) { a.why = std::string("nvjpeg symbol missing: ") + sym; return; }
    LUMI_SYM(CreateSimple, "nvjpegCreateSimple")
    LUMI_SYM(Destroy, "nvjpegDestroy")
    LUMI_SYM(JpegStateCreate, "nvjpegJpegStateCreate")
    LUMI_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy")
    LUMI_SYM(GetImageInfo, "nvjpegGetImageInfo")
    LUMI_SYM(Decode, "nvjpegDecode")
#undef LUMI_SYM
    a.ok = true;
  });
  return a;
}
}  // namespace

extern "C" {

const char* lumi_jpeg_last_error(void) { return g_jpeg_error.c_str(); }

int lumi_decode_jpeg(const unsigned char* data, size_t nbytes, int device, unsigned char* out, size_t capacity,
                     int out_on_device, int* height, int* width) {
  try {
    LUMI_REQUIRE(data && nbytes > 0 && height && width, "lumi_decode_jpeg: bad arguments");
    NvJpegApi& a = api();
    if (!a.ok) throw lumi::Error(LUMI_ECUDA, "nvJPEG unavailable: " + a.why);
    LUMI_CUDA_CHECK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(a.mu);
    if (!a.handle) {
      if (a.CreateSimple(&a.handle) != NVJPEG_STATUS_SUCCESS) throw lumi::Error(LUMI_ECUDA, "nvjpegCreateSimple failed");
      if (a.JpegStateCreate(a.handle, &a.state) != NVJPEG_STATUS_SUCCESS)
        throw lumi::Error(LUMI_ECUDA, "nvjpegJpegStateCreate failed");
    }
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    if (a.GetImageInfo(a.handle, data, nbytes, &ncomp, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS)
      throw lumi::Error(LUMI_EINVAL, "not a decodable JPEG stream");
    *height = hs[0]; *width = ws[0];
    if (!out) return LUMI_OK;                            // size query
    const size_t need = (size_t)hs[0] * ws[0] * 3;
    LUMI_REQUIRE(capacity >= need, "lumi_decode_jpeg: output buffer too small");
    unsigned char* dev = out;
    if (!out_on_device) LUMI_CUDA_CHECK(cudaMalloc(&dev, need));
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = dev;
    img.pitch[0] = (size_t)ws[0] * 3;
    const nvjpegStatus_t st = a.Decode(a.handle, a.state, data, nbytes, NVJPEG_OUTPUT_RGBI, &img, nullptr);
    cudaError_t ce = cudaStreamSynchronize(nullptr);
    if (st == NVJPEG_STATUS_SUCCESS && ce == cudaSuccess && !out_on_device)
      ce = cudaMemcpy(out, dev, need, cudaMemcpyDeviceToHost);
    if (!out_on_device) cudaFree(dev);
    if (st != NVJPEG_STATUS_SUCCESS) throw lumi::Error(LUMI_EINVAL, "nvjpegDecode failed (status " + std::to_string((int)st) + ")");
    LUMI_CUDA_CHECK(ce);
    return LUMI_OK;
  } catch (const lumi::Error& e) {
    g_jpeg_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_jpeg_error = e.what();
    return LUMI_EINVAL;
  }
}

}  // extern "C"
