// Count ingest in front of the fit: sparse layers -> the dense padded
// (cells, ld) float32 tiles the step kernel streams, without a dense host copy.
//
// The reference densifies every sparse layer on the host
// (brie/models/model_wrap.py:108-111, brie/models/model_TFProb.py:135-137), applies the
// pseudo-count with boolean-mask indexing (model_wrap.py:113-117) and builds two dense
// float64 (cells, genes) matrices for the gene filter (brie/utils/preprocessing.py:39-45).
// Here the CSC/CSR triplets cross PCIe (8 B per stored count instead of 4 B per matrix
// element and layer) and the scatter, pseudo-count and per-gene filter statistics run on
// the device.
#include <stdint.h>

#include "brie_host.h"

namespace brie {

// One warp per event column of a CSC slab; lanes stride over the column's stored counts.
// atomicAdd so that duplicate (row, col) entries sum, as scipy's toarray() does.
__global__ void __launch_bounds__(256) ingest_csc_kernel(const int64_t* __restrict__ colptr,
                                                         const int32_t* __restrict__ rows,
                                                         const float* __restrict__ vals, int64_t n_events,
                                                         int64_t n_cells, int64_t ld, float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_events; g += warps) {
    const int64_t lo = colptr[g], hi = colptr[g + 1];
    for (int64_t e = lo + lane; e < hi; e += 32) {
      const int64_t r = rows[e];
      if (r >= 0 && r < n_cells) atomicAdd(out + r * ld + g, vals[e]);
    }
  }
}

// One warp per cell row of a CSR matrix; keeps the columns of [event_begin, event_begin + n_events).
__global__ void __launch_bounds__(256) ingest_csr_kernel(const int64_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ cols,
                                                         const float* __restrict__ vals, int64_t n_cells,
                                                         int64_t event_begin, int64_t n_events, int64_t ld,
                                                         float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_cells; r += warps) {
    const int64_t lo = rowptr[r], hi = rowptr[r + 1];
    for (int64_t e = lo + lane; e < hi; e += 32) {
      const int64_t g = (int64_t)cols[e] - event_begin;
      if (g >= 0 && g < n_events) atomicAdd(out + r * ld + g, vals[e]);
    }
  }
}

// model_wrap.py:113-117: idx = c1 + c2 > 0; c1[idx] += pseudo; c2[idx] += pseudo   (float32)
__global__ void __launch_bounds__(256) pseudo_count_kernel(float4* c1, float4* c2, int64_t n4, float pseudo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = c1[i], b = c2[i];
    bool any = false;
    if (a.x + b.x > 0.f) { a.x += pseudo; b.x += pseudo; any = true; }
    if (a.y + b.y > 0.f) { a.y += pseudo; b.y += pseudo; any = true; }
    if (a.z + b.z > 0.f) { a.z += pseudo; b.z += pseudo; any = true; }
    if (a.w + b.w > 0.f) { a.w += pseudo; b.w += pseudo; any = true; }
    if (any) { c1[i] = a; c2[i] = b; }
  }
}

// Per-event filter statistics over a dense tile (preprocessing.py:38-61): column sums of each
// layer and the number of cells with unique (c1 + c2 > 0) / any (c1 + c2 + c3 > 0) counts.
// Stage 1: one thread per event, a CTA covers 128 events x one row chunk (coalesced 512-byte
// row segments); stage 2 adds the chunk partials in fixed order -> deterministic.
constexpr int kStatFields = 5;

__global__ void __launch_bounds__(128) gene_stats_partial_kernel(const float* __restrict__ c1,
                                                                 const float* __restrict__ c2,
                                                                 const float* __restrict__ c3, int64_t n_cells,
                                                                 int64_t ld, int64_t rows_per_chunk,
                                                                 double* part /* (chunks, 5, ld) */) {
  const int64_t g = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= ld) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(r0 + rows_per_chunk, n_cells);
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t nu = 0, nt = 0;
  for (int64_t r = r0; r < r1; ++r) {
    const float a = c1[r * ld + g], b = c2[r * ld + g], c = c3 ? c3[r * ld + g] : 0.f;
    s1 += (double)a; s2 += (double)b; s3 += (double)c;
    const double u = (double)a + (double)b;    // the reference accumulates the layers in float64
    nu += u > 0.0;
    nt += (u + (double)c) > 0.0;
  }
  double* p = part + (int64_t)blockIdx.y * kStatFields * ld + g;
  p[0] = s1; p[ld] = s2; p[2 * ld] = s3; p[3 * ld] = (double)nu; p[4 * ld] = (double)nt;
}

__global__ void __launch_bounds__(256) gene_stats_reduce_kernel(const double* __restrict__ part, int n_chunks,
                                                                int64_t ld, double* out /* (5, ld) */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over 5 * ld
  if (i >= kStatFields * ld) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += part[(int64_t)c * kStatFields * ld + i];
  out[i] = s;
}

// keep[j] -> column gather of a dense tile (events that pass the filter), so the kept counts
// never go back to the host: out[r, j] = in[r, src[j]].
__global__ void __launch_bounds__(256) gather_columns_kernel(const float* __restrict__ in, int64_t ld_in,
                                                             const int64_t* __restrict__ src, int64_t n_cells,
                                                             int64_t n_out, int64_t ld_out, float* out) {
  const int64_t total = n_cells * ld_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_out, j = i % ld_out;
    out[i] = j < n_out ? in[r * ld_in + src[j]] : 0.f;
  }
}

// inverse of the gather: out[r, dst[j]] = in[r, j]; columns not listed keep their contents
__global__ void __launch_bounds__(256) scatter_columns_kernel(const float* __restrict__ in, int64_t ld_in,
                                                              const int64_t* __restrict__ dst, int64_t n_cells,
                                                              int64_t n_in, int64_t ld_out, float* out) {
  const int64_t total = n_cells * n_in;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n_in, j = i % n_in;
    out[r * ld_out + dst[j]] = in[r * ld_in + j];
  }
}

namespace {
int stats_chunks(int64_t n_cells, int64_t ld, int64_t* rows_per_chunk) {
  const int64_t tiles = ceil_div(ld, 128);
  int64_t chunks = ceil_div((int64_t)148 * 16, tiles);
  if (chunks > n_cells) chunks = n_cells;
  if (chunks > 4096) chunks = 4096;
  if (chunks < 1) chunks = 1;
  *rows_per_chunk = ceil_div(n_cells, chunks);
  return (int)ceil_div(n_cells, *rows_per_chunk);
}
}  // namespace

}  // namespace brie

using namespace brie;

extern "C" {

int brie_ingest_csc(int64_t n_cells, int64_t n_events, int64_t ld, const int64_t* colptr, const int32_t* rows,
                    const float* vals, float* out, void* stream) {
  if (n_cells <= 0 || n_events <= 0 || ld < n_events) return fail(BRIE_ERR_ARG, "bad shape");
  if (!colptr || !out) return fail(BRIE_ERR_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  BRIE_CUDA(cudaMemsetAsync(out, 0, (size_t)n_cells * ld * sizeof(float), s));
  if (rows && vals) {
    ingest_csc_kernel<<<grid_1d(n_events * 32, 256), 256, 0, s>>>(colptr, rows, vals, n_events, n_cells, ld, out);
    BRIE_CUDA(cudaGetLastError());
  }
  return BRIE_OK;
}

int brie_ingest_csr(int64_t n_cells, int64_t event_begin, int64_t n_events, int64_t ld, const int64_t* rowptr,
                    const int32_t* cols, const float* vals, float* out, void* stream) {
  if (n_cells <= 0 || n_events <= 0 || ld < n_events || event_begin < 0) return fail(BRIE_ERR_ARG, "bad shape");
  if (!rowptr || !out) return fail(BRIE_ERR_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  BRIE_CUDA(cudaMemsetAsync(out, 0, (size_t)n_cells * ld * sizeof(float), s));
  if (cols && vals) {
    ingest_csr_kernel<<<grid_1d(n_cells * 32, 256), 256, 0, s>>>(rowptr, cols, vals, n_cells, event_begin, n_events,
                                                                ld, out);
    BRIE_CUDA(cudaGetLastError());
  }
  return BRIE_OK;
}

int brie_add_pseudo_count(int64_t n_cells, int64_t ld, float pseudo_count, float* c1, float* c2, void* stream) {
  if (n_cells <= 0 || ld <= 0 || ld % 4 != 0) return fail(BRIE_ERR_ARG, "bad shape (ld must be a multiple of 4)");
  if (!c1 || !c2) return fail(BRIE_ERR_ARG, "null argument");
  if ((((uintptr_t)c1) | ((uintptr_t)c2)) & 15u) return fail(BRIE_ERR_ARG, "buffers must be 16-byte aligned");
  const int64_t n4 = n_cells * ld / 4;
  pseudo_count_kernel<<<grid_1d(n4, 256), 256, 0, (cudaStream_t)stream>>>((float4*)c1, (float4*)c2, n4, pseudo_count);
  BRIE_CUDA(cudaGetLastError());
  return BRIE_OK;
}

size_t brie_gene_stats_scratch_bytes(int64_t n_cells, int64_t ld) {
  if (n_cells <= 0 || ld <= 0) return 0;
  int64_t rows;
  const int chunks = stats_chunks(n_cells, ld, &rows);
  return (size_t)chunks * kStatFields * ld * sizeof(double);
}

int brie_gene_stats(int64_t n_cells, int64_t ld, const float* c1, const float* c2, const float* c3, double* stats,
                    void* scratch, void* stream) {
  if (n_cells <= 0 || ld <= 0) return fail(BRIE_ERR_ARG, "bad shape");
  if (!c1 || !c2 || !stats || !scratch) return fail(BRIE_ERR_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t rows;
  const int chunks = stats_chunks(n_cells, ld, &rows);
  const dim3 grid((unsigned)ceil_div(ld, 128), (unsigned)chunks);
  gene_stats_partial_kernel<<<grid, 128, 0, s>>>(c1, c2, c3, n_cells, ld, rows, (double*)scratch);
  BRIE_CUDA(cudaGetLastError());
  gene_stats_reduce_kernel<<<(unsigned)ceil_div(kStatFields * ld, 256), 256, 0, s>>>((const double*)scratch, chunks,
                                                                                      ld, stats);
  BRIE_CUDA(cudaGetLastError());
  return BRIE_OK;
}

int brie_gather_events(int64_t n_cells, int64_t ld_in, const float* in, const int64_t* src, int64_t n_out,
                       int64_t ld_out, float* out, void* stream) {
  if (n_cells <= 0 || ld_in <= 0 || n_out < 0 || ld_out < n_out || ld_out <= 0) return fail(BRIE_ERR_ARG, "bad shape");
  if (!in || !out || (n_out > 0 && !src)) return fail(BRIE_ERR_ARG, "null argument");
  gather_columns_kernel<<<grid_1d(n_cells * ld_out, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, src, n_cells,
                                                                                         n_out, ld_out, out);
  BRIE_CUDA(cudaGetLastError());
  return BRIE_OK;
}

int brie_scatter_events(int64_t n_cells, int64_t ld_in, const float* in, const int64_t* dst, int64_t n_in,
                        int64_t ld_out, float* out, void* stream) {
  if (n_cells <= 0 || ld_in <= 0 || n_in < 0 || ld_in < n_in || ld_out <= 0) return fail(BRIE_ERR_ARG, "bad shape");
  if (!in || !out || (n_in > 0 && !dst)) return fail(BRIE_ERR_ARG, "null argument");
  if (n_in == 0) return BRIE_OK;
  scatter_columns_kernel<<<grid_1d(n_cells * n_in, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, dst, n_cells,
                                                                                        n_in, ld_out, out);
  BRIE_CUDA(cudaGetLastError());
  return BRIE_OK;
}

}  // extern "C"
