// merge_api.cu -- the final merge of a multi-GPU search as a one-sided push over NVLink: rank 0
// owns one buffer for the merged table (cudaMalloc, exported through CUDA IPC), every other rank
// maps it and writes its own guide rows straight into its slice -- all peers at once, no
// receive calls, no staging -- and rank 0 then computes the first-seen bucket ids in place
// (hawk_first_seen_dev). The reference has no counterpart (one process); row order and bucket
// ids are those of search_guides.py:306-337, :530-547 over the concatenated haplotype list.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "hawk_host.h"

namespace hawk {
__global__ void push_hap_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int64_t n, int32_t add,
                                int32_t keep /* index that is not shifted (REF), -1: none */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t h = src[i];
    dst[i] = h == keep ? h : h + add;
  }
}
}  // namespace hawk

static inline int64_t al256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int hawk_merge_layout(int64_t total_rows, int32_t text_stride, int32_t with_text, int64_t* off /* [7] */,
                                 int64_t* bytes) {
  if (total_rows < 0 || !off || !bytes || text_stride < 0) return hawk_fail(HAWK_EINVAL, "hawk_merge_layout: bad arguments");
  // hap, strand, pos, start, stop, bucket, text -- the column order of hawk_result_device_columns
  const int64_t w[7] = {4, 1, 4, 4, 4, 4, with_text ? text_stride : 0};
  int64_t at = 0;
  for (int k = 0; k < 7; ++k) {
    off[k] = at;
    at += al256(total_rows * w[k]);
  }
  *bytes = at ? at : 256;
  return HAWK_OK;
}

extern "C" int hawk_peer_alloc(hawk_ctx* c, int64_t bytes, void** d_ptr, uint8_t* handle /* [64] */) {
  if (!c || bytes <= 0 || !d_ptr || !handle) return hawk_fail(HAWK_EINVAL, "hawk_peer_alloc: bad arguments");
  CKCUDA(cudaSetDevice(c->device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) {  // give the context's cached blocks back and try once more
    cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    c->trim();
    CKCUDA(cudaMalloc(&p, (size_t)bytes));
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return hawk_check_cuda(e, "cudaIpcGetMemHandle");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle, &h, 64);
  *d_ptr = p;
  return HAWK_OK;
}

extern "C" int hawk_peer_free(hawk_ctx* c, void* d_ptr) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_peer_free: null context");
  if (!d_ptr) return HAWK_OK;
  CKCUDA(cudaSetDevice(c->device));
  CKCUDA(cudaFree(d_ptr));
  return HAWK_OK;
}

extern "C" int hawk_peer_open(hawk_ctx* c, const uint8_t* handle, void** d_ptr) {
  if (!c || !handle || !d_ptr) return hawk_fail(HAWK_EINVAL, "hawk_peer_open: bad arguments");
  CKCUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CKCUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HAWK_OK;
}

extern "C" int hawk_peer_close(hawk_ctx* c, void* d_ptr) {
  if (!c) return hawk_fail(HAWK_EINVAL, "hawk_peer_close: null context");
  if (!d_ptr) return HAWK_OK;
  CKCUDA(cudaSetDevice(c->device));
  CKCUDA(cudaStreamSynchronize(c->stream));
  CKCUDA(cudaIpcCloseMemHandle(d_ptr));
  return HAWK_OK;
}

extern "C" int hawk_result_push(hawk_result* r, int64_t row_lo, int32_t hap_add, int32_t hap_keep, void* d_buffer,
                                int64_t total_rows, int64_t at, int32_t with_text, int64_t* pushed_bytes) {
  if (!r || !d_buffer || row_lo < 0 || row_lo > r->n_guides || at < 0 || at + (r->n_guides - row_lo) > total_rows)
    return hawk_fail(HAWK_EINVAL, "hawk_result_push: bad arguments");
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t m = r->n_guides - row_lo;
  int64_t off[7], bytes = 0;
  CK(hawk_merge_layout(total_rows, r->text_stride, with_text, off, &bytes));
  if (pushed_bytes) *pushed_bytes = 0;
  if (m == 0) return HAWK_OK;
  uint8_t* base = (uint8_t*)d_buffer;
  // haplotype indices: local -> global on the way (peer stores over NVLink)
  {
    int64_t blocks = (m + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    hawk::push_hap_kernel<<<(unsigned)blocks, 256, 0, st>>>((int32_t*)(base + off[0]) + at, r->hap.as<int32_t>() + row_lo, m,
                                                            hap_add, hap_keep);
    hawk_note_launch(1);
    CK(hawk_check_cuda(cudaGetLastError(), "push_hap_kernel launch"));
  }
  auto copy = [&](int k, const void* src, int64_t width) -> int {
    return hawk_check_cuda(cudaMemcpyAsync(base + off[k] + at * width, (const uint8_t*)src + row_lo * width,
                                           (size_t)(m * width), cudaMemcpyDeviceToDevice, st), "peer copy");
  };
  CK(copy(1, r->strand.p, 1));
  CK(copy(2, r->pos.p, 4));
  CK(copy(3, r->start.p, 4));
  CK(copy(4, r->stop.p, 4));
  if (with_text) CK(copy(6, r->text.p, r->text_stride));
  if (pushed_bytes) *pushed_bytes = m * (17 + (with_text ? r->text_stride : 0));
  return HAWK_OK;
}
