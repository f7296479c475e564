// The inverted index behind the text clauses: CSR postings built from host arrays (rass_bm25_build*) or grown on the
// device from the token streams of bulk-indexed rows (rass_text_add_rows / rass_text_commit).
//
// Replaces what OpenSearch does with the `text` / `keyword` fields of a bulk request (reference app/main.py:1258-1269,
// mapping :361-561): Lucene inverts every refresh's documents into a segment (term -> ascending doc ids with term
// frequencies, plus one norm byte per document and field) and merges segments in the background.  Here a bulk's tokens
// become a segment on the device -- one stable radix sort by term of (term, row) keys (the rows arrive in ascending order,
// so a term's run is already doc-ordered) and a run-length pass that folds repeats into term frequencies -- and a commit
// folds the pending segments into the CSR the hybrid kernels walk: rows only ever grow, so a term's merged list is its old
// list followed by each segment's list, i.e. one copy pass over the postings, no comparison merge.  Lengths, norm bytes,
// docCount / sumTotalTermFreq, the per-tile offset table and the per-term score ranges are recomputed on the device as
// well; only V-sized statistics (df, idf) cross to the host.  Rewriting an indexed row is the one thing this does not
// do: the caller rebuilds from host arrays (rass_bm25_build_fields), as before.
//
// Sort, scan and run-length encode are CUB device primitives (the CUDA toolkit's own header library): index
// construction, not the query hot path.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

// ---- Lucene SmallFloat.intToByte4 / byte4ToInt (restated from the published algorithm) --------------------
__host__ __device__ static inline int long_to_int4(int64_t v) {
  int nb = 0;
  for (uint64_t t = (uint64_t)v; t; t >>= 1) ++nb;
  if (nb < 4) return (int)v;
  int shift = nb - 4;
  int enc = (int)((v >> shift) & 7);
  enc |= (shift + 1) << 3;
  return enc;
}
static int64_t int4_to_long(int e) {
  int bits = e & 7, shift = (e >> 3) - 1;
  return shift == -1 ? bits : (int64_t)(bits | 8) << shift;
}
#define RASS_NUM_FREE 24  // 255 - longToInt4(Integer.MAX_VALUE)
__host__ __device__ static inline uint8_t int_to_byte4(uint32_t i) {
  return (uint8_t)(i < (uint32_t)RASS_NUM_FREE ? i : RASS_NUM_FREE + long_to_int4((int64_t)i - RASS_NUM_FREE));
}
static int64_t byte4_to_int(int b) { return b < RASS_NUM_FREE ? b : RASS_NUM_FREE + int4_to_long(b - RASS_NUM_FREE); }

// Per-tile posting offsets of the frequent terms, built once per build / commit:
// tile_off[row * (n_tiles + 1) + t] = number of postings of the term whose doc is < t * HYB_TILE.
__global__ void tile_offsets_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ doc,
                                    const int32_t* __restrict__ table_terms, int n_table, int n_tiles,
                                    uint32_t* __restrict__ tile_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_table * (n_tiles + 1)) return;
  const int row = (int)(i / (n_tiles + 1)), t = (int)(i % (n_tiles + 1));
  const int32_t term = table_terms[row];
  const int64_t lo = indptr[term], hi = indptr[term + 1];
  const int64_t bound = (int64_t)t * HYB_TILE;
  int64_t a = lo, b = hi;                 // first posting with doc >= bound
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if ((int64_t)doc[m] < bound) a = m + 1; else b = m;
  }
  tile_off[i] = (uint32_t)(a - lo);
}

// Per-term range of x = tf * inv[norm] over the term's postings (one CTA per term), computed with the scoring kernel's
// own float ops.  s(x) = w - w / (1 + x) is non-decreasing in x under round-to-nearest, so s(xmin) / s(xmax) bound every
// score the term can contribute -- what hybrid_core needs to decide whether a query's clause sums are exact in double
// whatever the order of the additions (see hybrid_tile_fast_kernel).
__global__ void __launch_bounds__(256) term_xrange_kernel(const int64_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ doc,
                                                          const uint16_t* __restrict__ tf,
                                                          const uint8_t* __restrict__ norm, const float* __restrict__ inv,
                                                          const uint8_t* __restrict__ term_field, int64_t norm_rows,
                                                          float* __restrict__ xmin, float* __restrict__ xmax,
                                                          float* __restrict__ xq) {
  const int64_t t = blockIdx.x;
  const int64_t lo = indptr[t], hi = indptr[t + 1];
  const int f = term_field[t];
  const uint8_t* nf = norm + (size_t)f * norm_rows;
  const float* iv = inv + f * 256;
  float mn = __int_as_float(0x7f800000), mx = 0.f;
  for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
    const float x = __fmul_rn((float)tf[p], iv[nf[doc[p]]]);
    xq[p] = x;                        // the posting's impact, what hybrid_tile_fast_kernel scores from
    mn = fminf(mn, x);
    mx = fmaxf(mx, x);
  }
  __shared__ float smn[8], smx[8];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    xmin[t] = mn;
    xmax[t] = mx;
  }
}

// norm bytes of the [F][N] length planes + per-field docCount / sumTotalTermFreq
// (omit_mask bit f: the field omits norms -- a `keyword` field: every document that has a value counts as length 1)
__global__ void __launch_bounds__(256) norm_stats_kernel(const uint32_t* __restrict__ doclen, int64_t stride, int64_t N,
                                                         uint8_t* __restrict__ norm,
                                                         unsigned long long* __restrict__ stats,
                                                         const uint8_t* __restrict__ omit) {
  const int f = blockIdx.y;
  const bool om = omit[f] != 0;
  unsigned long long dc = 0, ttf = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t l = doclen[(size_t)f * stride + i];
    if (om) l = l != 0;
    norm[(size_t)f * N + i] = int_to_byte4(l);
    dc += l != 0;
    ttf += l;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    dc += __shfl_xor_sync(0xffffffffu, dc, m);
    ttf += __shfl_xor_sync(0xffffffffu, ttf, m);
  }
  if ((threadIdx.x & 31) == 0 && (dc | ttf)) {
    atomicAdd(stats + 2 * f, dc);
    atomicAdd(stats + 2 * f + 1, ttf);
  }
}

// RASS_DEBUG_TEXT_TIMES: wall-clock marks of the ingest / commit phases on stderr (measurement switch)
#include <chrono>
struct TextTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  const char* what;
  explicit TextTimer(const char* w) : on(getenv("RASS_DEBUG_TEXT_TIMES") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
  void mark(const char* label) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[text times] %s: %s %.2f ms\n", what, label, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

#define TEXT_TRY(h, call)                                                                                     \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess)                                                                                    \
      return rass_fail((h), e_ == cudaErrorMemoryAllocation ? RASS_E_OOM : RASS_E_CUDA, "%s failed: %s (%s:%d)", \
                       #call, cudaGetErrorString(e_), __FILE__, __LINE__);                                    \
  } while (0)

template <typename T>
static int upload(rass_engine* h, T** dst, const T* src, size_t n) {
  cudaFree(*dst);
  *dst = nullptr;
  TEXT_TRY(h, cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) TEXT_TRY(h, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return RASS_OK;
}

// the [F][stride] token-count planes on the device hold at least F fields x N rows (zero = the row lacks the field)
static int ensure_doclen(rass_engine* h, int F, int64_t N) {
  Bm25State& b = h->bm25;
  if (F <= b.doclen_F && N <= b.doclen_stride) return RASS_OK;
  const int nF = std::max(F, b.doclen_F);
  int64_t stride = b.doclen_stride;
  if (N > stride) stride = std::max<int64_t>(N, stride + stride / 2);
  stride = std::max<int64_t>(stride, 1024);
  uint32_t *p = nullptr, *g = nullptr;
  TEXT_TRY(h, cudaMalloc(&p, (size_t)nF * stride * sizeof(uint32_t)));
  TEXT_TRY(h, cudaMemset(p, 0, (size_t)nF * stride * sizeof(uint32_t)));
  TEXT_TRY(h, cudaMalloc(&g, (size_t)nF * stride * sizeof(uint32_t)));
  TEXT_TRY(h, cudaMemset(g, 0, (size_t)nF * stride * sizeof(uint32_t)));
  if (b.doclen_dev && b.doclen_F > 0 && b.doclen_stride > 0) {
    TEXT_TRY(h, cudaMemcpy2D(p, (size_t)stride * 4, b.doclen_dev, (size_t)b.doclen_stride * 4, (size_t)b.doclen_stride * 4,
                             (size_t)b.doclen_F, cudaMemcpyDeviceToDevice));
    TEXT_TRY(h, cudaMemcpy2D(g, (size_t)stride * 4, b.gen_dev, (size_t)b.doclen_stride * 4, (size_t)b.doclen_stride * 4,
                             (size_t)b.doclen_F, cudaMemcpyDeviceToDevice));
  }
  cudaFree(b.doclen_dev);
  cudaFree(b.gen_dev);
  b.doclen_dev = p;
  b.gen_dev = g;
  b.doclen_F = nF;
  b.doclen_stride = stride;
  return RASS_OK;
}

// Everything derived from the CSR (b.indptr / b.doc / b.tf on the device, b.indptr_host) and the length planes:
// norm bytes, per-field statistics and length tables, idf, the tile-offset table, the per-term score ranges.
// norm bytes of the [F][N] planes (b.norm) + per-field docCount / sumTotalTermFreq of THIS handle's rows
static int text_norms_and_stats(rass_engine* h, int64_t N, int F, std::vector<unsigned long long>* stats_out) {
  Bm25State& b = h->bm25;
  cudaFree(b.norm); b.norm = nullptr;
  TEXT_TRY(h, cudaMalloc(&b.norm, std::max<size_t>((size_t)N * F, 1)));
  std::vector<unsigned long long> stats((size_t)2 * F, 0);
  if (N > 0) {
    unsigned long long* stats_dev = nullptr;
    uint8_t* omit_dev = nullptr;
    std::vector<uint8_t> omit((size_t)F, 0);
    for (int f = 0; f < F && f < (int)b.field_omit_norms.size(); ++f) omit[(size_t)f] = b.field_omit_norms[(size_t)f];
    TEXT_TRY(h, cudaMalloc(&stats_dev, stats.size() * 8));
    cudaError_t e = cudaMalloc(&omit_dev, (size_t)F);
    if (e == cudaSuccess) e = cudaMemcpy(omit_dev, omit.data(), (size_t)F, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(stats_dev, 0, stats.size() * 8);
    if (e == cudaSuccess) {
      const dim3 grid((unsigned)std::min<int64_t>((N + 255) / 256, 1024), (unsigned)F);
      norm_stats_kernel<<<grid, 256>>>(b.doclen_dev, b.doclen_stride, N, b.norm, stats_dev, omit_dev);
      e = cudaMemcpy(stats.data(), stats_dev, stats.size() * 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(stats_dev);
    cudaFree(omit_dev);
    if (e != cudaSuccess) return rass_fail(h, RASS_E_CUDA, "norm_stats_kernel: %s", cudaGetErrorString(e));
  }
  *stats_out = stats;
  return RASS_OK;
}

// what a sharded handle sums over its shards between the merge and the finalisation of a commit
int text_local_stats(rass_engine* h, int F, int64_t N, int64_t* doc_count, int64_t* sum_ttf) {
  cudaSetDevice(h->device);
  std::vector<unsigned long long> stats;
  const int rc = text_norms_and_stats(h, N, F, &stats);
  if (rc) return rc;
  for (int f = 0; f < F; ++f) { doc_count[f] = (int64_t)stats[(size_t)2 * f]; sum_ttf[f] = (int64_t)stats[(size_t)2 * f + 1]; }
  return RASS_OK;
}

static int text_finalize(rass_engine* h, int64_t V, int64_t N, int F, const int64_t* g_doc_count,
                         const int64_t* g_sum_ttf, const int64_t* global_df) {
  Bm25State& b = h->bm25;
  const int64_t* indptr = b.indptr_host.data();
  b.V = V; b.N = N; b.nnz = indptr[V]; b.F = F;
  int rc;
  // norm planes + statistics
  TextTimer tt("finalize");
  std::vector<unsigned long long> stats;
  if ((rc = text_norms_and_stats(h, N, F, &stats))) return rc;
  tt.mark("norms + stats");
  std::vector<float> inv((size_t)256 * F);
  std::vector<int64_t> doc_count((size_t)F, 0);
  const float k1 = 1.2f, bb = 0.75f, one = 1.0f;
  for (int f = 0; f < F; ++f) {
    int64_t dc = (int64_t)stats[(size_t)2 * f], sum_ttf = (int64_t)stats[(size_t)2 * f + 1];
    if (g_doc_count && g_sum_ttf && g_doc_count[f] > 0) { dc = g_doc_count[f]; sum_ttf = g_sum_ttf[f]; }
    doc_count[(size_t)f] = dc;
    if (b.field_doc_count.size() < (size_t)F) { b.field_doc_count.resize((size_t)F, 0); b.field_sum_ttf.resize((size_t)F, 0); }
    b.field_doc_count[(size_t)f] = dc;
    b.field_sum_ttf[(size_t)f] = sum_ttf;
    const float avgdl = dc ? (float)((double)sum_ttf / (double)dc) : 0.f;
    if (f == 0) { b.doc_count = dc; b.avgdl = avgdl; }
    for (int i = 0; i < 256; ++i) {
      if (!dc) { inv[(size_t)f * 256 + i] = 0.f; continue; }
      volatile float t = bb * (float)byte4_to_int(i);   // volatile: every step rounds to float, no contraction
      t = t / avgdl;
      t = (one - bb) + t;
      t = k1 * t;
      inv[(size_t)f * 256 + i] = one / t;
    }
  }
  b.idf_host.resize((size_t)V);
  for (int64_t t = 0; t < V; ++t) {
    const int f = b.term_field_host[(size_t)t];
    const int64_t df = global_df ? global_df[t] : indptr[t + 1] - indptr[t];
    const double dc = (double)doc_count[(size_t)f];
    b.idf_host[(size_t)t] = (float)log(1.0 + (dc - (double)df + 0.5) / ((double)df + 0.5));
  }
  if ((rc = upload(h, &b.inv_dev, inv.data(), inv.size()))) return rc;
  tt.mark("idf + length tables");
  // per-tile posting offsets of the frequent terms (the tile kernels jump straight to a tile's postings)
  b.n_tiles = (int)((std::max<int64_t>(N, 1) + HYB_TILE - 1) / HYB_TILE);
  b.table_row_host.assign((size_t)V, -1);
  std::vector<int32_t> table_terms;
  for (int64_t t = 0; t < V; ++t)
    if (indptr[t + 1] - indptr[t] >= HYB_TABLE_MIN_DF) {
      b.table_row_host[(size_t)t] = (int32_t)table_terms.size();
      table_terms.push_back((int32_t)t);
    }
  cudaFree(b.tile_off); b.tile_off = nullptr;
  const size_t n_off = table_terms.size() * (size_t)(b.n_tiles + 1);
  TEXT_TRY(h, cudaMalloc(&b.tile_off, std::max<size_t>(n_off, 1) * sizeof(uint32_t)));
  if (n_off) {
    int32_t* tt_dev = nullptr;
    if ((rc = upload(h, &tt_dev, table_terms.data(), table_terms.size()))) return rc;
    tile_offsets_kernel<<<(unsigned)((n_off + 255) / 256), 256>>>(b.indptr, b.doc, tt_dev, (int)table_terms.size(),
                                                                  b.n_tiles, b.tile_off);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(tt_dev);
    if (e != cudaSuccess) return rass_fail(h, RASS_E_CUDA, "tile_offsets_kernel: %s", cudaGetErrorString(e));
  }
  tt.mark("tile offsets");
  // per-term score ranges (order-free fast path of hybrid_core)
  b.xmin_host.assign((size_t)V, 0.f);
  b.xmax_host.assign((size_t)V, 0.f);
  if (V > 0) {
    uint8_t* tfield_dev = nullptr;
    float *xmin_dev = nullptr, *xmax_dev = nullptr;
    if ((rc = upload(h, &tfield_dev, b.term_field_host.data(), (size_t)V))) return rc;
    cudaFree(b.xq); b.xq = nullptr;
    cudaError_t e = cudaMalloc(&xmin_dev, (size_t)V * 4);
    if (e == cudaSuccess) e = cudaMalloc(&xmax_dev, (size_t)V * 4);
    if (e == cudaSuccess) e = cudaMalloc(&b.xq, std::max<size_t>((size_t)b.nnz, 1) * 4);
    if (e == cudaSuccess) {
      term_xrange_kernel<<<(unsigned)V, 256>>>(b.indptr, b.doc, b.tf, b.norm, b.inv_dev, tfield_dev, N, xmin_dev, xmax_dev,
                                              b.xq);
      e = cudaMemcpy(b.xmin_host.data(), xmin_dev, (size_t)V * 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(b.xmax_host.data(), xmax_dev, (size_t)V * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(tfield_dev); cudaFree(xmin_dev); cudaFree(xmax_dev);
    if (e != cudaSuccess) return rass_fail(h, RASS_E_CUDA, "term_xrange_kernel: %s", cudaGetErrorString(e));
  }
  tt.mark("impacts + term ranges");
  b.built = true;
  return RASS_OK;
}

static void drop_segments(Bm25State& b) {
  for (TextSegment& s : b.pending) { cudaFree(s.uterm); cudaFree(s.uptr); cudaFree(s.doc); cudaFree(s.tf); }
  b.pending.clear();
}

// F analysed fields share one CSR: term t belongs to field term_field[t]; doclen is [F][N] (tokens of the field per
// row, 0 = the row does not have the field).  Statistics (docCount, avgdl, idf) are per field, as in Lucene.
int bm25_build_impl(rass_engine* h, const int64_t* indptr, const int32_t* doc, const uint16_t* tf,
                    const int32_t* term_field, const uint32_t* doclen, int64_t V, int64_t N, int F,
                    const int64_t* g_doc_count, const int64_t* g_sum_ttf, const int64_t* global_df) {
  if (!h) return RASS_E_INVALID;
  cudaSetDevice(h->device);
  if (V < 0 || N < 0 || F < 1 || F > 255 || !indptr || (N && !doclen)) return rass_fail(h, RASS_E_INVALID, "bad postings");
  const int64_t nnz = indptr[V];
  if (nnz < 0 || (nnz && (!doc || !tf))) return rass_fail(h, RASS_E_INVALID, "bad postings");
  if (N > 0x7ffffff0LL) return rass_fail(h, RASS_E_INVALID, "too many documents");
  if (term_field)
    for (int64_t t = 0; t < V; ++t)
      if (term_field[t] < 0 || term_field[t] >= F) return rass_fail(h, RASS_E_INVALID, "term %lld: bad field", (long long)t);
  Bm25State& b = h->bm25;
  drop_segments(b);
  b.indptr_host.assign(indptr, indptr + V + 1);
  b.term_field_host.assign((size_t)V, 0);
  // a later rass_text_add_rows continues from this index when every field's terms are one contiguous block, in field
  // order (what the client builds); the last row each field holds is read off the length planes
  b.field_vocab.assign((size_t)F, 0);
  b.contiguous_fields = true;
  for (int64_t t = 0; t < V; ++t) {
    const int f = term_field ? term_field[t] : 0;
    b.term_field_host[(size_t)t] = (uint8_t)f;
    if (t > 0 && f < b.term_field_host[(size_t)t - 1]) b.contiguous_fields = false;
    b.field_vocab[(size_t)f]++;
  }
  b.field_last_row.assign((size_t)F, -1);
  for (int f = 0; f < F; ++f)
    for (int64_t i = N - 1; i >= 0; --i)
      if (doclen[(size_t)f * N + i]) { b.field_last_row[(size_t)f] = i; break; }
  int rc;
  if ((rc = upload(h, &b.indptr, indptr, (size_t)V + 1))) return rc;
  if ((rc = upload(h, &b.doc, doc, (size_t)nnz))) return rc;
  if ((rc = upload(h, &b.tf, tf, (size_t)nnz))) return rc;
  cudaFree(b.doclen_dev); b.doclen_dev = nullptr; b.doclen_F = 0; b.doclen_stride = 0;
  cudaFree(b.gen_dev); b.gen_dev = nullptr;
  b.has_rewrite = false;
  if ((rc = ensure_doclen(h, F, N))) return rc;
  if (N)
    TEXT_TRY(h, cudaMemcpy2D(b.doclen_dev, (size_t)b.doclen_stride * 4, doclen, (size_t)N * 4, (size_t)N * 4, (size_t)F,
                             cudaMemcpyHostToDevice));
  return text_finalize(h, V, N, F, g_doc_count, g_sum_ttf, global_df);
}

extern "C" int rass_bm25_build(rass_engine* h, const int64_t* indptr, const int32_t* doc, const uint16_t* tf,
                               const uint32_t* doclen, int64_t V, int64_t N, int64_t global_doc_count,
                               int64_t global_sum_ttf, const int64_t* global_df) {
  SHARDED(h, sharded_bm25_build(h, indptr, doc, tf, nullptr, doclen, V, N, 1));
  const bool global = global_doc_count > 0;
  return bm25_build_impl(h, indptr, doc, tf, nullptr, doclen, V, N, 1, global ? &global_doc_count : nullptr,
                         global ? &global_sum_ttf : nullptr, global_df);
}

extern "C" int rass_bm25_build_fields(rass_engine* h, const int64_t* indptr, const int32_t* doc, const uint16_t* tf,
                                      const int32_t* term_field, const uint32_t* doclen, int64_t V, int64_t N, int F) {
  SHARDED(h, sharded_bm25_build(h, indptr, doc, tf, term_field, doclen, V, N, F));
  if (h && !term_field) return rass_fail(h, RASS_E_INVALID, "null term_field");
  return bm25_build_impl(h, indptr, doc, tf, term_field, doclen, V, N, F, nullptr, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// device-side ingest: a bulk's token stream -> a segment
// ---------------------------------------------------------------------------------------------
// key = term << 32 | position of the row inside the bulk
__global__ void token_keys_kernel(const int64_t* __restrict__ tok_indptr, const int32_t* __restrict__ tok_terms,
                                  int64_t n_rows, uint64_t* __restrict__ keys, int* __restrict__ bad) {
  const int64_t r = blockIdx.x;                      // one CTA per row: its tokens are contiguous
  const int64_t lo = tok_indptr[r], hi = tok_indptr[r + 1];
  for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const int32_t t = tok_terms[j];
    if (t < 0) atomicExch(bad, 1);
    keys[j] = ((uint64_t)(uint32_t)t << 32) | (uint64_t)r;
  }
}

// the lengths of the bulk's rows go into the field's plane (once the token stream has been validated)
// ... and the segment's number into its generation plane: a posting of a source (0 = the committed CSR, s = the s-th
// pending segment) is live iff the generation of its (field, doc) still names that source
__global__ void doclen_scatter_kernel(const int64_t* __restrict__ tok_indptr, const int64_t* __restrict__ rows,
                                      int64_t n_rows, uint32_t* __restrict__ doclen_plane,
                                      uint32_t* __restrict__ gen_plane, uint32_t seg_id) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) {
    doclen_plane[rows[r]] = (uint32_t)(tok_indptr[r + 1] - tok_indptr[r]);
    gen_plane[rows[r]] = seg_id;
  }
}

// unique (term, row position) keys + repeat counts -> doc ids, term frequencies, "first posting of a new term" marks
__global__ void split_runs_kernel(const uint64_t* __restrict__ ukeys, const int* __restrict__ counts, int64_t n,
                                  const int64_t* __restrict__ rows, int32_t* __restrict__ doc, uint16_t* __restrict__ tf,
                                  int32_t* __restrict__ term_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = ukeys[i];
  doc[i] = (int32_t)rows[k & 0xffffffffu];
  tf[i] = (uint16_t)min(counts[i], 65535);
  term_of[i] = (int32_t)(k >> 32);
}

// scratch of the segment builds: grow-only between commits (a cudaMalloc / cudaFree pair per buffer and bulk costs
// more than the sort), released by rass_text_commit
static int ensure_scratch(rass_engine* h, size_t bytes) {
  Bm25State& b = h->bm25;
  if (bytes <= b.scratch_bytes) return RASS_OK;
  cudaFree(b.scratch);
  b.scratch = nullptr;
  b.scratch_bytes = 0;
  bytes += bytes / 4;
  TEXT_TRY(h, cudaMalloc(&b.scratch, bytes));
  b.scratch_bytes = bytes;
  return RASS_OK;
}

static int seg_build(rass_engine* h, int field, const int64_t* rows_dev, int64_t n_rows, const int64_t* indptr_dev,
                     const int32_t* terms_dev, int64_t T, int32_t max_term, cudaStream_t st) {
  Bm25State& b = h->bm25;
  TextSegment seg;
  seg.field = field;
  const uint32_t seg_id = (uint32_t)b.pending.size() + 1;
  if (T == 0) {                  // rows without tokens: lengths 0, whatever they held before is dropped at the commit
    doclen_scatter_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(
        indptr_dev, rows_dev, n_rows, b.doclen_dev + (size_t)field * b.doclen_stride,
        b.gen_dev + (size_t)field * b.doclen_stride, seg_id);
    TEXT_TRY(h, cudaGetLastError());
    b.pending.push_back(seg);
    return RASS_OK;
  }
  int term_bits = 1;
  while (term_bits < 31 && ((int64_t)1 << term_bits) <= (int64_t)max_term) ++term_bits;
  auto done = [&](int code) {
    if (code) { cudaFree(seg.uterm); cudaFree(seg.uptr); cudaFree(seg.doc); cudaFree(seg.tf); }
    return code;
  };
#define SEG_TRY(call)                                                                                          \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return done(rass_fail(h, e_ == cudaErrorMemoryAllocation ? RASS_E_OOM : RASS_E_CUDA, "%s failed: %s (%s:%d)", \
                            #call, cudaGetErrorString(e_), __FILE__, __LINE__));                               \
  } while (0)
  // temporary storage of the four CUB calls (host-side size queries), then one arena:
  // [k0 u64 x T][k1 u64 x T][cnt i32 x T][term_of i32 x T][uterm i32 x T][n_runs, bad][cub temp]
  size_t tmp_bytes = 0;
  {
    cub::DoubleBuffer<uint64_t> kq(nullptr, nullptr);
    size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    SEG_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t1, kq, (int)T, 32, 32 + term_bits, st));
    SEG_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, t2, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr,
                                               (int*)nullptr, (int)T, st));
    SEG_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, t3, (int32_t*)nullptr, (int32_t*)nullptr, (int*)nullptr,
                                               (int*)nullptr, (int)T, st));
    SEG_TRY(cub::DeviceScan::InclusiveSum(nullptr, t4, (int*)nullptr, (int64_t*)nullptr, (int)T, st));
    tmp_bytes = std::max(std::max(t1, t2), std::max(t3, t4));
  }
  const size_t Tp = ((size_t)T + 63) & ~(size_t)63;      // keeps every sub-array 256-byte aligned
  int rc;
  if ((rc = ensure_scratch(h, Tp * 28 + 256 + tmp_bytes + 256))) return done(rc);
  unsigned char* base = static_cast<unsigned char*>(b.scratch);
  uint64_t* k0 = reinterpret_cast<uint64_t*>(base);
  uint64_t* k1 = reinterpret_cast<uint64_t*>(base + Tp * 8);
  int* cnt = reinterpret_cast<int*>(base + Tp * 16);
  int32_t* term_of = reinterpret_cast<int32_t*>(base + Tp * 20);
  int32_t* uterm_tmp = reinterpret_cast<int32_t*>(base + Tp * 24);
  int* n_runs_dev = reinterpret_cast<int*>(base + Tp * 28);
  int* bad_dev = n_runs_dev + 1;
  void* tmp = base + Tp * 28 + 256;
  SEG_TRY(cudaMemsetAsync(bad_dev, 0, sizeof(int), st));
  token_keys_kernel<<<(unsigned)n_rows, 128, 0, st>>>(indptr_dev, terms_dev, n_rows, k0, bad_dev);
  SEG_TRY(cudaGetLastError());
  // stable sort by the term bits only: a term's run keeps the bulk's row order, which is ascending
  cub::DoubleBuffer<uint64_t> keys(k0, k1);
  SEG_TRY(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, (int)T, 32, 32 + term_bits, st));
  // repeats of (term, row) fold into a term frequency
  uint64_t* sorted = keys.Current();
  uint64_t* uk = keys.Alternate();            // the other buffer is free again: the unique keys go there
  SEG_TRY(cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, sorted, uk, cnt, n_runs_dev, (int)T, st));
  int n_post = 0, bad = 0;
  SEG_TRY(cudaMemcpyAsync(&n_post, n_runs_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  SEG_TRY(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  SEG_TRY(cudaStreamSynchronize(st));
  if (bad) return done(rass_fail(h, RASS_E_INVALID, "negative term id in the token stream"));
  doclen_scatter_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(
      indptr_dev, rows_dev, n_rows, b.doclen_dev + (size_t)field * b.doclen_stride,
      b.gen_dev + (size_t)field * b.doclen_stride, seg_id);
  seg.n_post = n_post;
  SEG_TRY(cudaMalloc(&seg.doc, (size_t)n_post * 4));
  SEG_TRY(cudaMalloc(&seg.tf, (size_t)n_post * 2));
  split_runs_kernel<<<(unsigned)((n_post + 255) / 256), 256, 0, st>>>(uk, cnt, n_post, rows_dev, seg.doc, seg.tf, term_of);
  SEG_TRY(cudaGetLastError());
  // the segment's own term directory: distinct terms + where each one's run starts
  int* ucnt = cnt;                             // reuse: run lengths of the term sequence
  SEG_TRY(cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, term_of, uterm_tmp, ucnt, n_runs_dev, n_post, st));
  int n_uniq = 0;
  SEG_TRY(cudaMemcpyAsync(&n_uniq, n_runs_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  SEG_TRY(cudaStreamSynchronize(st));
  seg.n_uniq = n_uniq;
  SEG_TRY(cudaMalloc(&seg.uterm, std::max<size_t>((size_t)n_uniq, 1) * 4));
  SEG_TRY(cudaMemcpyAsync(seg.uterm, uterm_tmp, (size_t)n_uniq * 4, cudaMemcpyDeviceToDevice, st));
  SEG_TRY(cudaMalloc(&seg.uptr, ((size_t)n_uniq + 1) * 8));
  SEG_TRY(cudaMemsetAsync(seg.uptr, 0, 8, st));
  SEG_TRY(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, ucnt, seg.uptr + 1, n_uniq, st));
  SEG_TRY(cudaStreamSynchronize(st));
#undef SEG_TRY
  b.pending.push_back(seg);
  return done(RASS_OK);
}

static int text_add_rows_impl(rass_engine* h, int field, const int64_t* rows, int64_t n_rows, const int64_t* tok_indptr,
                              const int32_t* tok_terms, bool on_device) {
  if (!h) return RASS_E_INVALID;
  if (h->shards)
    return rass_fail(h, RASS_E_UNSUPPORTED, "a handle over several GPUs takes token streams from host memory (rass_text_add_rows)");
  cudaSetDevice(h->device);
  Bm25State& b = h->bm25;
  if (field < 0 || field > 254) return rass_fail(h, RASS_E_INVALID, "bad field %d", field);
  if (n_rows < 0 || (n_rows && (!rows || !tok_indptr))) return rass_fail(h, RASS_E_INVALID, "bad token stream");
  if (n_rows == 0) return RASS_OK;
  if (b.built && !b.contiguous_fields)
    return rass_fail(h, RASS_E_UNSUPPORTED, "the index was built with interleaved field terms: rebuild from host arrays");
  cudaStream_t st = eng_stream(h);
  // host copies of the offsets and the row ids: validated here, they size everything
  std::vector<int64_t> ip((size_t)n_rows + 1), rw((size_t)n_rows);
  if (on_device) {
    TEXT_TRY(h, cudaMemcpyAsync(ip.data(), tok_indptr, ip.size() * 8, cudaMemcpyDeviceToHost, st));
    TEXT_TRY(h, cudaMemcpyAsync(rw.data(), rows, rw.size() * 8, cudaMemcpyDeviceToHost, st));
    TEXT_TRY(h, cudaStreamSynchronize(st));
  } else {
    memcpy(ip.data(), tok_indptr, ip.size() * 8);
    memcpy(rw.data(), rows, rw.size() * 8);
  }
  const int64_t T = ip[(size_t)n_rows] - ip[0];
  if (ip[0] != 0 || T < 0 || T > 0x7fffffffLL) return rass_fail(h, RASS_E_INVALID, "a bulk holds at most 2^31 - 1 tokens");
  if (T && !tok_terms) return rass_fail(h, RASS_E_INVALID, "bad token stream");
  const int64_t last = (size_t)field < b.field_last_row.size() ? b.field_last_row[(size_t)field] : -1;
  for (int64_t i = 0; i < n_rows; ++i) {
    if (ip[(size_t)i + 1] < ip[(size_t)i]) return rass_fail(h, RASS_E_INVALID, "token offsets must not decrease");
    if (rw[(size_t)i] < 0 || (i && rw[(size_t)i] <= rw[(size_t)i - 1]))
      return rass_fail(h, RASS_E_INVALID, "the rows of a bulk must be ascending and distinct (row %lld after %lld)",
                       (long long)rw[(size_t)i], (long long)(i ? rw[(size_t)i - 1] : -1));
  }
  if (rw[(size_t)n_rows - 1] > 0x7ffffff0LL) return rass_fail(h, RASS_E_INVALID, "too many documents");
  // a row at or below one the field already holds is a REWRITE: the commit drops what the row held before (its
  // generation no longer names the old source) and re-sorts instead of concatenating
  const bool rewrites = rw[0] <= last;
  int rc;
  if ((rc = ensure_doclen(h, std::max(field + 1, std::max(b.F, b.doclen_F)), rw[(size_t)n_rows - 1] + 1))) return rc;
  int64_t *rows_dev = nullptr, *ip_dev = nullptr;
  int32_t* terms_dev = nullptr;
  const int64_t* rows_d = rows;
  const int64_t* ip_d = tok_indptr;
  const int32_t* terms_d = tok_terms;
  int32_t max_term = 0;
  auto cleanup = [&]() { cudaFree(rows_dev); cudaFree(ip_dev); cudaFree(terms_dev); };
  if (!on_device) {
    cudaError_t e = cudaMalloc(&rows_dev, (size_t)n_rows * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ip_dev, ((size_t)n_rows + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&terms_dev, std::max<size_t>((size_t)T, 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rows_dev, rows, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ip_dev, tok_indptr, ((size_t)n_rows + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && T) e = cudaMemcpyAsync(terms_dev, tok_terms, (size_t)T * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      cleanup();
      return rass_fail(h, e == cudaErrorMemoryAllocation ? RASS_E_OOM : RASS_E_CUDA, "text ingest staging: %s",
                       cudaGetErrorString(e));
    }
    rows_d = rows_dev; ip_d = ip_dev; terms_d = terms_dev;
    for (int64_t j = 0; j < T; ++j) max_term = std::max(max_term, tok_terms[j]);
  } else {
    max_term = 0x7fffffff;       // not known without a pass over the tokens: sort on all 31 bits
  }
  rc = seg_build(h, field, rows_d, n_rows, ip_d, terms_d, T, max_term, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) return rass_fail(h, RASS_E_CUDA, "text ingest: %s", cudaGetErrorString(e));
  if (b.field_last_row.size() <= (size_t)field) b.field_last_row.resize((size_t)field + 1, -1);
  b.field_last_row[(size_t)field] = std::max(b.field_last_row[(size_t)field], rw[(size_t)n_rows - 1]);
  if (rewrites) b.has_rewrite = true;
  return RASS_OK;
}

extern "C" int rass_text_add_rows(rass_engine* h, int field, const int64_t* rows, int64_t n_rows,
                                  const int64_t* tok_indptr, const int32_t* tok_terms) {
  SHARDED(h, sharded_text_add_rows(h, field, rows, n_rows, tok_indptr, tok_terms));
  return text_add_rows_impl(h, field, rows, n_rows, tok_indptr, tok_terms, false);
}

extern "C" int rass_text_add_rows_dev(rass_engine* h, int field, const int64_t* rows_dev, int64_t n_rows,
                                      const int64_t* tok_indptr_dev, const int32_t* tok_terms_dev) {
  return text_add_rows_impl(h, field, rows_dev, n_rows, tok_indptr_dev, tok_terms_dev, true);
}

// ---------------------------------------------------------------------------------------------
// commit: old CSR + pending segments -> new CSR
// ---------------------------------------------------------------------------------------------
__global__ void old_df_kernel(const int64_t* __restrict__ old_indptr, const int32_t* __restrict__ remap, int64_t V_old,
                              int64_t* __restrict__ df_new) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < V_old) df_new[remap[t]] = old_indptr[t + 1] - old_indptr[t];
}

__global__ void seg_df_kernel(const int32_t* __restrict__ uterm, const int64_t* __restrict__ uptr, int64_t n_uniq,
                              int64_t base, int64_t* __restrict__ df_new) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_uniq) df_new[base + uterm[u]] += uptr[u + 1] - uptr[u];
}

__global__ void copy_old_kernel(const int64_t* __restrict__ old_indptr, const int32_t* __restrict__ remap, int64_t V_old,
                                int64_t nnz_old, const int32_t* __restrict__ doc_old, const uint16_t* __restrict__ tf_old,
                                const int64_t* __restrict__ new_indptr, int32_t* __restrict__ doc_new,
                                uint16_t* __restrict__ tf_new) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz_old) return;
  int64_t a = 0, b = V_old;                    // last term whose list starts at or before p
  while (b - a > 1) {
    const int64_t m = (a + b) >> 1;
    if (old_indptr[m] <= p) a = m; else b = m;
  }
  const int64_t dst = new_indptr[remap[a]] + (p - old_indptr[a]);
  doc_new[dst] = doc_old[p];
  tf_new[dst] = tf_old[p];
}

__global__ void cursor_init_kernel(const int64_t* __restrict__ new_indptr, const int64_t* __restrict__ old_indptr,
                                   const int32_t* __restrict__ remap, int64_t V_old, int64_t V_new,
                                   int64_t* __restrict__ cursor) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < V_new) cursor[t] = new_indptr[t];
}
__global__ void cursor_old_kernel(const int64_t* __restrict__ old_indptr, const int32_t* __restrict__ remap, int64_t V_old,
                                  int64_t* __restrict__ cursor) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < V_old) cursor[remap[t]] += old_indptr[t + 1] - old_indptr[t];
}

__global__ void copy_seg_kernel(const int32_t* __restrict__ uterm, const int64_t* __restrict__ uptr, int64_t n_uniq,
                                int64_t n_post, int64_t base, const int32_t* __restrict__ doc, const uint16_t* __restrict__ tf,
                                const int64_t* __restrict__ cursor, int32_t* __restrict__ doc_new,
                                uint16_t* __restrict__ tf_new) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_post) return;
  int64_t a = 0, b = n_uniq;
  while (b - a > 1) {
    const int64_t m = (a + b) >> 1;
    if (uptr[m] <= j) a = m; else b = m;
  }
  const int64_t dst = cursor[base + uterm[a]] + (j - uptr[a]);
  doc_new[dst] = doc[j];
  tf_new[dst] = tf[j];
}

__global__ void cursor_seg_kernel(const int32_t* __restrict__ uterm, const int64_t* __restrict__ uptr, int64_t n_uniq,
                                  int64_t base, int64_t* __restrict__ cursor) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_uniq) cursor[base + uterm[u]] += uptr[u + 1] - uptr[u];
}

// ---- the re-sort commit: some pending segment rewrites rows the index already held ----
// Every source's live postings (generation of (field, doc) == the source) become keys `new term << db | doc`, dead ones
// the sentinel `V_new << db`; one radix sort of (key, tf) pairs IS the new CSR order.
__global__ void gather_old_kernel(const int64_t* __restrict__ old_indptr, const int32_t* __restrict__ remap,
                                  const uint8_t* __restrict__ tfield_old, int64_t V_old, int64_t nnz_old,
                                  const int32_t* __restrict__ doc_old, const uint16_t* __restrict__ tf_old,
                                  const uint32_t* __restrict__ gen, int64_t stride, int db, uint64_t sentinel,
                                  uint64_t* __restrict__ keys, uint16_t* __restrict__ vals,
                                  unsigned long long* __restrict__ n_live) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (p < nnz_old) {
    int64_t a = 0, b = V_old;
    while (b - a > 1) {
      const int64_t m = (a + b) >> 1;
      if (old_indptr[m] <= p) a = m; else b = m;
    }
    const int32_t d = doc_old[p];
    live = gen[(size_t)tfield_old[a] * stride + d] == 0u;
    keys[p] = live ? (((uint64_t)(uint32_t)remap[a] << db) | (uint64_t)(uint32_t)d) : sentinel;
    vals[p] = tf_old[p];
  }
  const unsigned bal = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_live, (unsigned long long)__popc(bal));
}

__global__ void gather_seg_kernel(const int32_t* __restrict__ uterm, const int64_t* __restrict__ uptr, int64_t n_uniq,
                                  int64_t n_post, int64_t base, const int32_t* __restrict__ doc,
                                  const uint16_t* __restrict__ tf, const uint32_t* __restrict__ gen_plane, uint32_t seg_id,
                                  int db, uint64_t sentinel, uint64_t* __restrict__ keys, uint16_t* __restrict__ vals,
                                  unsigned long long* __restrict__ n_live) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (j < n_post) {
    int64_t a = 0, b = n_uniq;
    while (b - a > 1) {
      const int64_t m = (a + b) >> 1;
      if (uptr[m] <= j) a = m; else b = m;
    }
    const int32_t d = doc[j];
    live = gen_plane[d] == seg_id;
    keys[j] = live ? (((uint64_t)(base + uterm[a]) << db) | (uint64_t)(uint32_t)d) : sentinel;
    vals[j] = tf[j];
  }
  const unsigned bal = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_live, (unsigned long long)__popc(bal));
}

__global__ void split_sorted_kernel(const uint64_t* __restrict__ keys, const uint16_t* __restrict__ vals, int64_t n,
                                    uint64_t doc_mask, int32_t* __restrict__ doc, uint16_t* __restrict__ tf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { doc[i] = (int32_t)(keys[i] & doc_mask); tf[i] = vals[i]; }
}

// indptr[t] = number of sorted keys below `t << db`
__global__ void term_bounds_kernel(const uint64_t* __restrict__ keys, int64_t n, int db, int64_t V,
                                   int64_t* __restrict__ indptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > V) return;
  const uint64_t want = (uint64_t)t << db;
  int64_t a = 0, b = n;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (keys[m] < want) a = m + 1; else b = m;
  }
  indptr[t] = a;
}

static int commit_resort(rass_engine* h, const std::vector<int64_t>& base_new, const int32_t* remap_dev, int64_t V_old,
                         int64_t nnz_old, int64_t V_new, int64_t N, cudaStream_t st, int64_t** indptr_out,
                         int32_t** doc_out, uint16_t** tf_out) {
  Bm25State& b = h->bm25;
  int64_t total = nnz_old;
  for (const TextSegment& s : b.pending) total += s.n_post;
  if (total > 0x7fffffffLL) return rass_fail(h, RASS_E_UNSUPPORTED, "a re-sorting commit handles at most 2^31 - 1 postings");
  int db = 1, tb = 1;
  while (db < 32 && ((int64_t)1 << db) < std::max<int64_t>(N, 2)) ++db;
  while (tb < 32 && ((int64_t)1 << tb) <= V_new) ++tb;          // the sentinel term V_new must be representable
  const uint64_t sentinel = (uint64_t)V_new << db;
  uint64_t *k0 = nullptr, *k1 = nullptr;
  uint16_t *v0 = nullptr, *v1 = nullptr;
  uint8_t* tfield_dev = nullptr;
  unsigned long long* n_live_dev = nullptr;
  void* tmp = nullptr;
  int64_t* indptr_new = nullptr;
  int32_t* doc_new = nullptr;
  uint16_t* tf_new = nullptr;
  auto done = [&](int code) {
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tfield_dev); cudaFree(n_live_dev); cudaFree(tmp);
    if (code) { cudaFree(indptr_new); cudaFree(doc_new); cudaFree(tf_new); }
    return code;
  };
#define RESORT_TRY(call)                                                                                       \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return done(rass_fail(h, e_ == cudaErrorMemoryAllocation ? RASS_E_OOM : RASS_E_CUDA, "%s failed: %s (%s:%d)", \
                            #call, cudaGetErrorString(e_), __FILE__, __LINE__));                               \
  } while (0)
  const size_t n = std::max<size_t>((size_t)total, 1);
  RESORT_TRY(cudaMalloc(&k0, n * 8));
  RESORT_TRY(cudaMalloc(&k1, n * 8));
  RESORT_TRY(cudaMalloc(&v0, n * 2));
  RESORT_TRY(cudaMalloc(&v1, n * 2));
  RESORT_TRY(cudaMalloc(&n_live_dev, 8));
  RESORT_TRY(cudaMemsetAsync(n_live_dev, 0, 8, st));
  RESORT_TRY(cudaMalloc(&tfield_dev, std::max<size_t>((size_t)V_old, 1)));
  if (V_old)
    RESORT_TRY(cudaMemcpyAsync(tfield_dev, b.term_field_host.data(), (size_t)V_old, cudaMemcpyHostToDevice, st));
  const unsigned TB = 256;
  auto blocks = [&](int64_t m) { return (unsigned)std::max<int64_t>((m + TB - 1) / TB, 1); };
  if (nnz_old)
    gather_old_kernel<<<blocks(nnz_old), TB, 0, st>>>(b.indptr, remap_dev, tfield_dev, V_old, nnz_old, b.doc, b.tf, b.gen_dev,
                                                     b.doclen_stride, db, sentinel, k0, v0, n_live_dev);
  int64_t off = nnz_old;
  for (size_t i = 0; i < b.pending.size(); ++i) {
    const TextSegment& s = b.pending[i];
    if (!s.n_post) continue;
    gather_seg_kernel<<<blocks(s.n_post), TB, 0, st>>>(s.uterm, s.uptr, s.n_uniq, s.n_post, base_new[(size_t)s.field], s.doc,
                                                      s.tf, b.gen_dev + (size_t)s.field * b.doclen_stride, (uint32_t)i + 1,
                                                      db, sentinel, k0 + off, v0 + off, n_live_dev);
    off += s.n_post;
  }
  RESORT_TRY(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> keys(k0, k1);
  cub::DoubleBuffer<uint16_t> vals(v0, v1);
  size_t tmp_bytes = 0;
  RESORT_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, vals, (int)total, 0, db + tb, st));
  RESORT_TRY(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  if (total) RESORT_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, vals, (int)total, 0, db + tb, st));
  unsigned long long n_live = 0;
  RESORT_TRY(cudaMemcpyAsync(&n_live, n_live_dev, 8, cudaMemcpyDeviceToHost, st));
  RESORT_TRY(cudaStreamSynchronize(st));
  RESORT_TRY(cudaMalloc(&indptr_new, ((size_t)V_new + 1) * 8));
  RESORT_TRY(cudaMalloc(&doc_new, std::max<size_t>((size_t)n_live, 1) * 4));
  RESORT_TRY(cudaMalloc(&tf_new, std::max<size_t>((size_t)n_live, 1) * 2));
  if (n_live)
    split_sorted_kernel<<<blocks((int64_t)n_live), TB, 0, st>>>(keys.Current(), vals.Current(), (int64_t)n_live,
                                                               ((uint64_t)1 << db) - 1, doc_new, tf_new);
  term_bounds_kernel<<<blocks(V_new + 1), TB, 0, st>>>(keys.Current(), (int64_t)n_live, db, V_new, indptr_new);
  RESORT_TRY(cudaGetLastError());
  b.indptr_host.assign((size_t)V_new + 1, 0);
  RESORT_TRY(cudaMemcpyAsync(b.indptr_host.data(), indptr_new, ((size_t)V_new + 1) * 8, cudaMemcpyDeviceToHost, st));
  RESORT_TRY(cudaStreamSynchronize(st));
#undef RESORT_TRY
  *indptr_out = indptr_new;
  *doc_out = doc_new;
  *tf_out = tf_new;
  return done(RASS_OK);
}

// The merge half of a commit: new CSR, length planes and b.indptr_host in place; nothing derived from corpus statistics yet
// (text_finalize / text_finalize_global does that, with this handle's statistics or a sharded handle's sums).
int text_commit_merge(rass_engine* h, const int64_t* field_vocab, int F, int64_t N) {
  if (!h) return RASS_E_INVALID;
  cudaSetDevice(h->device);
  Bm25State& b = h->bm25;
  if (!field_vocab || F < 1 || F > 255 || N < 0 || N > 0x7ffffff0LL) return rass_fail(h, RASS_E_INVALID, "bad commit");
  if (b.built && !b.contiguous_fields)
    return rass_fail(h, RASS_E_UNSUPPORTED, "the index was built with interleaved field terms: rebuild from host arrays");
  const int F_old = b.built ? b.F : 0;
  const int64_t V_old = b.built ? b.V : 0, nnz_old = b.built ? b.nnz : 0;
  if (F < F_old) return rass_fail(h, RASS_E_INVALID, "fields cannot disappear (%d < %d)", F, F_old);
  std::vector<int64_t> base_new((size_t)F + 1, 0);
  for (int f = 0; f < F; ++f) {
    const int64_t old_v = f < F_old ? b.field_vocab[(size_t)f] : 0;
    if (field_vocab[f] < old_v) return rass_fail(h, RASS_E_INVALID, "the vocabulary of field %d cannot shrink", f);
    base_new[(size_t)f + 1] = base_new[(size_t)f] + field_vocab[f];
  }
  const int64_t V_new = base_new[(size_t)F];
  if (V_new > 0x7ffffff0LL) return rass_fail(h, RASS_E_INVALID, "too many terms");
  for (const TextSegment& s : b.pending)
    if (s.field >= F) return rass_fail(h, RASS_E_INVALID, "a pending segment belongs to field %d of %d", s.field, F);
  for (size_t f = 0; f < b.field_last_row.size(); ++f)
    if (b.field_last_row[f] >= N) return rass_fail(h, RASS_E_INVALID, "field %d holds row %lld >= N", (int)f,
                                                   (long long)b.field_last_row[f]);
  int rc;
  TextTimer tt("commit");
  if ((rc = ensure_doclen(h, F, N))) return rc;
  cudaStream_t st = eng_stream(h);
  tt.mark("planes");
  // old term -> new term (the blocks of the fields move apart as vocabularies grow)
  std::vector<int32_t> remap((size_t)V_old);
  {
    int64_t t = 0;
    for (int f = 0; f < F_old; ++f)
      for (int64_t l = 0; l < b.field_vocab[(size_t)f]; ++l) remap[(size_t)t++] = (int32_t)(base_new[(size_t)f] + l);
  }
  int32_t* remap_dev = nullptr;
  int64_t *df_new = nullptr, *indptr_new = nullptr, *cursor = nullptr;
  int32_t* doc_new = nullptr;
  uint16_t* tf_new = nullptr;
  void* tmp = nullptr;
  auto done = [&](int code) {
    cudaFree(remap_dev); cudaFree(df_new); cudaFree(cursor); cudaFree(tmp);
    if (code) { cudaFree(indptr_new); cudaFree(doc_new); cudaFree(tf_new); }
    return code;
  };
#define COMMIT_TRY(call)                                                                                       \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return done(rass_fail(h, e_ == cudaErrorMemoryAllocation ? RASS_E_OOM : RASS_E_CUDA, "%s failed: %s (%s:%d)", \
                            #call, cudaGetErrorString(e_), __FILE__, __LINE__));                               \
  } while (0)
  const unsigned TB = 256;
  auto blocks = [&](int64_t n) { return (unsigned)std::max<int64_t>((n + TB - 1) / TB, 1); };
  COMMIT_TRY(cudaMalloc(&remap_dev, std::max<size_t>((size_t)V_old, 1) * 4));
  if (V_old) COMMIT_TRY(cudaMemcpyAsync(remap_dev, remap.data(), (size_t)V_old * 4, cudaMemcpyHostToDevice, st));
  if (b.has_rewrite) {
    // rows were rewritten: drop what they held, re-sort everything that is live
    if ((rc = commit_resort(h, base_new, remap_dev, V_old, nnz_old, V_new, N, st, &indptr_new, &doc_new, &tf_new)))
      return done(rc);
  } else {
  COMMIT_TRY(cudaMalloc(&df_new, ((size_t)V_new + 1) * 8));
  COMMIT_TRY(cudaMemsetAsync(df_new, 0, ((size_t)V_new + 1) * 8, st));
  if (V_old) old_df_kernel<<<blocks(V_old), TB, 0, st>>>(b.indptr, remap_dev, V_old, df_new);
  for (const TextSegment& s : b.pending)
    if (s.n_uniq)
      seg_df_kernel<<<blocks(s.n_uniq), TB, 0, st>>>(s.uterm, s.uptr, s.n_uniq, base_new[(size_t)s.field], df_new);
  COMMIT_TRY(cudaGetLastError());
  COMMIT_TRY(cudaMalloc(&indptr_new, ((size_t)V_new + 1) * 8));
  size_t tmp_bytes = 0;
  COMMIT_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, df_new, indptr_new, (int)(V_new + 1), st));
  COMMIT_TRY(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  COMMIT_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, df_new, indptr_new, (int)(V_new + 1), st));
  b.indptr_host.assign((size_t)V_new + 1, 0);
  COMMIT_TRY(cudaMemcpyAsync(b.indptr_host.data(), indptr_new, ((size_t)V_new + 1) * 8, cudaMemcpyDeviceToHost, st));
  COMMIT_TRY(cudaStreamSynchronize(st));
  const int64_t nnz_new = b.indptr_host[(size_t)V_new];
  COMMIT_TRY(cudaMalloc(&doc_new, std::max<size_t>((size_t)nnz_new, 1) * 4));
  COMMIT_TRY(cudaMalloc(&tf_new, std::max<size_t>((size_t)nnz_new, 1) * 2));
  COMMIT_TRY(cudaMalloc(&cursor, std::max<size_t>((size_t)V_new, 1) * 8));
  if (V_new) cursor_init_kernel<<<blocks(V_new), TB, 0, st>>>(indptr_new, b.indptr, remap_dev, V_old, V_new, cursor);
  if (nnz_old)
    copy_old_kernel<<<blocks(nnz_old), TB, 0, st>>>(b.indptr, remap_dev, V_old, nnz_old, b.doc, b.tf, indptr_new, doc_new,
                                                   tf_new);
  if (V_old) cursor_old_kernel<<<blocks(V_old), TB, 0, st>>>(b.indptr, remap_dev, V_old, cursor);
  for (const TextSegment& s : b.pending) {
    if (!s.n_post) continue;
    const int64_t base = base_new[(size_t)s.field];
    copy_seg_kernel<<<blocks(s.n_post), TB, 0, st>>>(s.uterm, s.uptr, s.n_uniq, s.n_post, base, s.doc, s.tf, cursor, doc_new,
                                                    tf_new);
    cursor_seg_kernel<<<blocks(s.n_uniq), TB, 0, st>>>(s.uterm, s.uptr, s.n_uniq, base, cursor);
  }
  COMMIT_TRY(cudaGetLastError());
  COMMIT_TRY(cudaStreamSynchronize(st));
  }
  tt.mark(b.has_rewrite ? "re-sort" : "merge copy");
  // every (field, doc) is the committed index's again
  if (b.gen_dev)
    COMMIT_TRY(cudaMemsetAsync(b.gen_dev, 0, (size_t)b.doclen_F * b.doclen_stride * sizeof(uint32_t), st));
  b.has_rewrite = false;
#undef COMMIT_TRY
  cudaFree(b.indptr); cudaFree(b.doc); cudaFree(b.tf);
  b.indptr = indptr_new; b.doc = doc_new; b.tf = tf_new;
  drop_segments(b);
  cudaFree(b.scratch); b.scratch = nullptr; b.scratch_bytes = 0;
  b.field_vocab.assign(field_vocab, field_vocab + F);
  b.contiguous_fields = true;
  b.term_field_host.assign((size_t)V_new, 0);
  for (int f = 0; f < F; ++f)
    std::fill(b.term_field_host.begin() + base_new[(size_t)f], b.term_field_host.begin() + base_new[(size_t)f + 1], (uint8_t)f);
  if (b.field_last_row.size() < (size_t)F) b.field_last_row.resize((size_t)F, -1);
  b.pending_V = V_new;
  tt.mark("swap in");
  return done(RASS_OK);
}

int text_finalize_global(rass_engine* h, int64_t N, int F, const int64_t* g_doc_count, const int64_t* g_sum_ttf,
                         const int64_t* global_df) {
  cudaSetDevice(h->device);
  return text_finalize(h, h->bm25.pending_V, N, F, g_doc_count, g_sum_ttf, global_df);
}

extern "C" int rass_text_commit(rass_engine* h, const int64_t* field_vocab, int F, int64_t N) {
  SHARDED(h, sharded_text_commit(h, field_vocab, F, N));
  const int rc = text_commit_merge(h, field_vocab, F, N);
  if (rc) return rc;
  return text_finalize(h, h->bm25.pending_V, N, F, nullptr, nullptr, nullptr);
}

extern "C" int rass_text_omit_norms(rass_engine* h, int field, int omit) {
  if (!h) return RASS_E_INVALID;
  if (field < 0 || field > 254) return rass_fail(h, RASS_E_INVALID, "bad field %d", field);
  if (h->shards) {
    for (rass_engine* sh : sharded_shards(h)) {
      const int rc = rass_text_omit_norms(sh, field, omit);
      if (rc) return rc;
    }
    return RASS_OK;
  }
  Bm25State& b = h->bm25;
  if (b.field_omit_norms.size() <= (size_t)field) b.field_omit_norms.resize((size_t)field + 1, 0);
  b.field_omit_norms[(size_t)field] = omit ? 1 : 0;
  return RASS_OK;
}

extern "C" int rass_text_size(rass_engine* h, int64_t* V, int64_t* N, int64_t* nnz, int* F) {
  if (!h) return RASS_E_INVALID;
  SHARDED(h, sharded_text_size(h, V, N, nnz, F));
  const Bm25State& b = h->bm25;
  if (V) *V = b.built ? b.V : 0;
  if (N) *N = b.built ? b.N : 0;
  if (nnz) *nnz = b.built ? b.nnz : 0;
  if (F) *F = b.built ? b.F : 0;
  return RASS_OK;
}

extern "C" int rass_text_stats(rass_engine* h, int64_t* indptr, int64_t* doc_count, int64_t* sum_ttf) {
  if (!h) return RASS_E_INVALID;
  SHARDED(h, sharded_text_stats(h, indptr, doc_count, sum_ttf));
  const Bm25State& b = h->bm25;
  if (!b.built) return rass_fail(h, RASS_E_NOTFOUND, "no postings");
  if (indptr) memcpy(indptr, b.indptr_host.data(), ((size_t)b.V + 1) * 8);
  for (int f = 0; f < b.F; ++f) {
    if (doc_count) doc_count[f] = b.field_doc_count[(size_t)f];
    if (sum_ttf) sum_ttf[f] = b.field_sum_ttf[(size_t)f];
  }
  return RASS_OK;
}

extern "C" int rass_text_export(rass_engine* h, int64_t* indptr, int32_t* doc, uint16_t* tf, uint32_t* doclen,
                                uint8_t* norm) {
  if (!h) return RASS_E_INVALID;
  if (h->shards) return rass_fail(h, RASS_E_UNSUPPORTED, "device-side text ingest runs on single-device handles");
  cudaSetDevice(h->device);
  const Bm25State& b = h->bm25;
  if (!b.built) return rass_fail(h, RASS_E_NOTFOUND, "no postings");
  TEXT_TRY(h, cudaStreamSynchronize(eng_stream(h)));
  if (indptr) memcpy(indptr, b.indptr_host.data(), ((size_t)b.V + 1) * 8);
  if (doc && b.nnz) TEXT_TRY(h, cudaMemcpy(doc, b.doc, (size_t)b.nnz * 4, cudaMemcpyDeviceToHost));
  if (tf && b.nnz) TEXT_TRY(h, cudaMemcpy(tf, b.tf, (size_t)b.nnz * 2, cudaMemcpyDeviceToHost));
  if (doclen && b.N)
    TEXT_TRY(h, cudaMemcpy2D(doclen, (size_t)b.N * 4, b.doclen_dev, (size_t)b.doclen_stride * 4, (size_t)b.N * 4, (size_t)b.F,
                             cudaMemcpyDeviceToHost));
  if (norm && b.N) TEXT_TRY(h, cudaMemcpy(norm, b.norm, (size_t)b.N * b.F, cudaMemcpyDeviceToHost));
  return RASS_OK;
}
