// featurize_api.cu -- N4, second half: the learned scorers' inputs for every guide row of a
// device-resident table, batched. The reference builds them guide by guide in Python:
//   scoring.py:50-84   _extract_guide_sequences / _extract_guide_sequences_sgdesigner --
//                      guide.sequence[(PAD - lead) : (-PAD + 3)].upper(), lead = 4 for Azimuth, RS3,
//                      DeepCpf1 and CRISPRon (30-mers for G + P = 23, 34-mers for Cpf1), 0 for
//                      sgDesigner; the sequence is the one annotation.reverse_guides left
//                      (strand 1: IUPAC-aware reverse complement, guide.py:245-255);
//   scores/deepCpf1/seqdeepcpf1.py:71-92   preprocess -- the float32 one-hot tensor
//                      [guides, 4 (A, C, G, T), letters] DeepCpf1's Conv1d reads (a per-letter
//                      Python double loop); any other letter is a KeyError there.
// The models themselves stay the reference's host code; this hands them their whole batch at once
// (and, for DeepCpf1, already as the tensor and already on the device if the caller wants it there).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hawk_core.h"
#include "hawk_kernels.h"
#include "hawk_host.h"

namespace hawk {

constexpr int FZ_ROWS = 128;  // rows per block = threads per block

// Block b handles rows [128 b, 128 b + 128): (0) the rows' raw window texts are staged in shared
// memory with 128-bit loads (rows are consecutive in HBM), (1) one thread per row writes the
// row's feature letters next to them, (2) all threads stream the letters / the one-hot floats out
// as one contiguous run per block.
__global__ void __launch_bounds__(FZ_ROWS) featurize_kernel(const uint8_t* __restrict__ strand,
                                                            const uint8_t* __restrict__ text, int32_t text_stride, int W,
                                                            int lead, int L, int64_t n, uint8_t* __restrict__ kmers,
                                                            float* __restrict__ onehot,
                                                            unsigned long long* __restrict__ bad) {
  extern __shared__ uint4 fz_smem[];
  uint8_t* raw = (uint8_t*)fz_smem;                          // FZ_ROWS x text_stride
  uint8_t* feat = raw + (size_t)FZ_ROWS * text_stride;       // FZ_ROWS x L
  const int64_t row0 = (int64_t)blockIdx.x * FZ_ROWS;
  const int rows = (int)((n - row0) < FZ_ROWS ? (n - row0) : FZ_ROWS);
  const int vec_per_row = text_stride >> 4;
  const uint4* src = (const uint4*)(text + row0 * text_stride);
  for (int i = threadIdx.x; i < rows * vec_per_row; i += FZ_ROWS) fz_smem[i] = __ldg(src + i);
  __syncthreads();
  if ((int)threadIdx.x < rows) {
    const int s = strand[row0 + threadIdx.x];
    const uint8_t* mine = raw + (size_t)threadIdx.x * text_stride;
    uint8_t* out = feat + (size_t)threadIdx.x * L;
    bool ok = true;
    for (int j = 0; j < L; ++j) {
      const uint8_t u = feature_byte(mine, W, s, lead, j);
      out[j] = u;
      ok &= onehot_channel(u) >= 0;
    }
    if (onehot && !ok) atomicMin(bad, (unsigned long long)(row0 + threadIdx.x));
  }
  __syncthreads();
  if (kmers) {
    uint8_t* dst = kmers + row0 * L;
    for (int i = threadIdx.x; i < rows * L; i += FZ_ROWS) dst[i] = feat[i];
  }
  if (onehot) {
    float* dst = onehot + row0 * 4 * L;
    const int per_row = 4 * L;
    // a thread keeps its (channel, letter) slot of the row and walks the rows: no division in the
    // inner loop, a warp writes one contiguous run of a row
    for (int e = threadIdx.x; e < per_row; e += FZ_ROWS) {
      const int ch = e / L, j = e - ch * L;
      for (int rr = 0; rr < rows; ++rr)
        dst[(size_t)rr * per_row + e] = onehot_channel(feat[rr * L + j]) == ch ? 1.0f : 0.0f;
    }
  }
}

}  // namespace hawk

using namespace hawk;

extern "C" int hawk_result_featurize(hawk_result* r, int32_t lead, uint8_t* kmers, float* onehot, int32_t onehot_on_device,
                                     int64_t* bad_row) {
  if (bad_row) *bad_row = -1;
  if (!r || lead < 0 || lead > HAWK_GUIDESEQPAD)
    return hawk_fail(HAWK_EINVAL, "hawk_result_featurize: needs a guide table and 0 <= lead <= %d", HAWK_GUIDESEQPAD);
  const int64_t n = r->n_guides;
  if (n == 0) return HAWK_OK;
  if (!r->text.p || r->text_stride <= 0 || (r->text_stride & 15) || !r->strand.p)
    return hawk_fail(HAWK_EINVAL, "hawk_result_featurize: the table holds no window text column");
  if (!kmers && !onehot) return hawk_fail(HAWK_EINVAL, "hawk_result_featurize: no output requested");
  hawk_ctx* c = r->ctx;
  CKCUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int W = r->window, L = feature_len(W, lead);
  DevBuf d_k, d_o, d_bad;
  if (kmers) CK(d_k.alloc(c, (size_t)n * L));
  float* o_dev = nullptr;
  if (onehot) {
    if (onehot_on_device) {
      o_dev = onehot;
    } else {
      CK(d_o.alloc(c, (size_t)n * 4 * L * sizeof(float)));
      o_dev = d_o.as<float>();
    }
  }
  const unsigned long long none = ~0ull;
  CK(upload(c, d_bad, &none, 8));
  const size_t smem = (size_t)FZ_ROWS * (r->text_stride + L);
  featurize_kernel<<<(unsigned)((n + FZ_ROWS - 1) / FZ_ROWS), FZ_ROWS, smem, st>>>(
      r->strand.as<uint8_t>(), r->text.as<uint8_t>(), r->text_stride, W, lead, L, n, d_k.as<uint8_t>(), o_dev,
      d_bad.as<unsigned long long>());
  hawk_note_launch(1);
  CK(hawk_check_cuda(cudaGetLastError(), "featurize_kernel launch"));
  if (kmers) {
    c->d2h_bytes += n * L;
    CKCUDA(cudaMemcpyAsync(kmers, d_k.p, (size_t)n * L, cudaMemcpyDeviceToHost, st));
  }
  if (onehot && !onehot_on_device) {
    c->d2h_bytes += n * 4 * L * (int64_t)sizeof(float);
    CKCUDA(cudaMemcpyAsync(onehot, d_o.p, (size_t)n * 4 * L * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  unsigned long long bad = none;
  c->d2h_bytes += 8;
  CKCUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, st));
  CKCUDA(cudaStreamSynchronize(st));
  if (bad != none) {
    if (bad_row) *bad_row = (int64_t)bad;
    return hawk_fail(HAWK_EFEATURE, "hawk_result_featurize: row %llu holds a letter other than A, C, G, T (the reference's "
                     "one-hot encoding raises KeyError there)", bad);
  }
  return HAWK_OK;
}
