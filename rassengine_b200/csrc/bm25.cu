// BM25 term scoring over CSR postings and the bool.should boosted-sum fusion with the kNN clause.
//
// Replaces the Lucene side of OpenSearchIndexer.hybrid_search (reference app/main.py:1574-1598):
//   multi_match(best_fields, operator or) over `unstructuredText`  -> BM25Similarity postings walk
//   knn clause                                                       -> only the k nearest docs score
//   bool.should                                                      -> sum of the matching clauses
// Arithmetic follows oracle/bm25.py and oracle/fusion.py operation for operation (float ops rounded
// individually, clause sums in double, final cast to float) so ranked ids and scores are identical.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

int search_core(rass_engine* h, const float* q_dev, int B, int k, int64_t* out_rows, float* out_scores,
                double* out_keys, rass_stats* stats);
int stage_queries(rass_engine* h, const float* q_host, int B, float** q_dev_out);

// ---- Lucene SmallFloat.intToByte4 / byte4ToInt (restated from the published algorithm) --------------------
static int long_to_int4(int64_t v) {
  int nb = v == 0 ? 0 : 64 - __builtin_clzll((unsigned long long)v);
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
static const int kNumFree = 24;  // 255 - longToInt4(Integer.MAX_VALUE)
static uint8_t int_to_byte4(uint32_t i) { return (uint8_t)(i < (uint32_t)kNumFree ? i : kNumFree + long_to_int4((int64_t)i - kNumFree)); }
static int64_t byte4_to_int(int b) { return b < kNumFree ? b : kNumFree + int4_to_long(b - kNumFree); }

template <typename T>
static int upload(rass_engine* h, T** dst, const T* src, size_t n) {
  cudaFree(*dst);
  *dst = nullptr;
  CUDA_TRY(h, cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CUDA_TRY(h, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return RASS_OK;
}

extern "C" int rass_bm25_build(rass_engine* h, const int64_t* indptr, const int32_t* doc, const uint16_t* tf,
                               const uint32_t* doclen, int64_t V, int64_t N, int64_t global_doc_count,
                               int64_t global_sum_ttf, const int64_t* global_df) {
  if (!h) return RASS_E_INVALID;
  cudaSetDevice(h->device);
  if (V < 0 || N < 0 || !indptr || (N && !doclen)) return rass_fail(h, RASS_E_INVALID, "bad postings");
  const int64_t nnz = indptr[V];
  if (nnz < 0 || (nnz && (!doc || !tf))) return rass_fail(h, RASS_E_INVALID, "bad postings");
  if (N > 0xfffffff0LL) return rass_fail(h, RASS_E_INVALID, "too many documents");
  Bm25State& b = h->bm25;
  b.V = V; b.N = N; b.nnz = nnz;
  int64_t doc_count = 0, sum_ttf = 0;
  std::vector<uint8_t> norm((size_t)N);
  for (int64_t i = 0; i < N; ++i) {
    doc_count += doclen[i] != 0;
    sum_ttf += doclen[i];
    norm[(size_t)i] = int_to_byte4(doclen[i]);
  }
  if (global_doc_count > 0) { doc_count = global_doc_count; sum_ttf = global_sum_ttf; }
  b.doc_count = doc_count;
  b.avgdl = doc_count ? (float)((double)sum_ttf / (double)doc_count) : 0.f;
  b.indptr_host.assign(indptr, indptr + V + 1);
  b.idf_host.resize((size_t)V);
  for (int64_t t = 0; t < V; ++t) {
    const int64_t df = global_df ? global_df[t] : indptr[t + 1] - indptr[t];
    b.idf_host[(size_t)t] = (float)log(1.0 + ((double)doc_count - (double)df + 0.5) / ((double)df + 0.5));
  }
  float inv[256];
  const float k1 = 1.2f, bb = 0.75f, one = 1.0f;
  for (int i = 0; i < 256; ++i) {
    if (!doc_count) { inv[i] = 0.f; continue; }
    volatile float t = bb * (float)byte4_to_int(i);   // volatile: every step rounds to float, no contraction
    t = t / b.avgdl;
    t = (one - bb) + t;
    t = k1 * t;
    inv[i] = one / t;
  }
  int rc;
  if ((rc = upload(h, &b.indptr, indptr, (size_t)V + 1))) return rc;
  if ((rc = upload(h, &b.doc, doc, (size_t)nnz))) return rc;
  if ((rc = upload(h, &b.tf, tf, (size_t)nnz))) return rc;
  if ((rc = upload(h, &b.norm, norm.data(), (size_t)N))) return rc;
  if ((rc = upload(h, &b.inv_dev, inv, (size_t)256))) return rc;
  cudaFree(b.acc); b.acc = nullptr;
  cudaFree(b.touched); b.touched = nullptr;
  cudaFree(b.touched_n); b.touched_n = nullptr;
  const size_t n_acc = (size_t)std::max<int64_t>(std::max<int64_t>(N, h->n_rows), 1);
  CUDA_TRY(h, cudaMalloc(&b.acc, n_acc * sizeof(double)));
  CUDA_TRY(h, cudaMemset(b.acc, 0, n_acc * sizeof(double)));
  b.acc_rows = (int64_t)n_acc;
  b.touched_cap = (int64_t)n_acc;
  CUDA_TRY(h, cudaMalloc(&b.touched, n_acc * sizeof(uint32_t)));
  CUDA_TRY(h, cudaMalloc(&b.touched_n, sizeof(int)));
  CUDA_TRY(h, cudaMemset(b.touched_n, 0, sizeof(int)));
  return RASS_OK;
}

// ---- kernels ------------------------------------------------------------------------------------------------
#define BM25_MAX_TERMS 64

struct TermArgs {
  int n_terms;
  int64_t lo[BM25_MAX_TERMS];      // first posting of the term
  int64_t cum[BM25_MAX_TERMS + 1]; // prefix sum of posting counts
  float w[BM25_MAX_TERMS];         // float(boost) * idf
};

// one thread per posting of the query's terms: s = w - w / (1 + tf * inv[norm[d]])  (float, each op rounded)
__global__ void __launch_bounds__(256) bm25_accumulate_kernel(const __grid_constant__ TermArgs ta,
                                                              const int32_t* __restrict__ doc,
                                                              const uint16_t* __restrict__ tf,
                                                              const uint8_t* __restrict__ norm,
                                                              const float* __restrict__ inv,
                                                              const float* __restrict__ sb, int64_t n_rows,
                                                              const uint8_t* __restrict__ row_filter,
                                                              int64_t filter_rows, double* __restrict__ acc,
                                                              uint32_t* __restrict__ touched,
                                                              int* __restrict__ touched_n) {
  __shared__ float s_inv[256];
  s_inv[threadIdx.x] = inv[threadIdx.x];
  __syncthreads();
  const int64_t total = ta.cum[ta.n_terms];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = 0;
    while (i >= ta.cum[t + 1]) ++t;
    const int64_t p = ta.lo[t] + (i - ta.cum[t]);
    const uint32_t d = (uint32_t)__ldg(doc + p);
    if (row_filter && ((int64_t)d >= filter_rows || !row_filter[d])) continue;   // bool.filter
    const float w = ta.w[t];
    const float x = __fmul_rn((float)__ldg(tf + p), s_inv[norm[d]]);
    const float s = __fsub_rn(w, __fdiv_rn(w, __fadd_rn(1.0f, x)));
    if (!(s > 0.f)) continue;   // a doc matches the clause only with a positive score (oracle: text > 0)
    const double old = atomicAdd(acc + d, (double)s);
    if (old == 0.0) touched[atomicAdd(touched_n, 1)] = d;
  }
}

// the knn clause: the k nearest rows of this query get float(w_knn * knn_score) added
__global__ void fuse_knn_kernel(const int64_t* __restrict__ knn_rows, const float* __restrict__ knn_scores, int k,
                                int64_t row_base, float w_knn, const uint8_t* __restrict__ row_filter,
                                int64_t filter_rows, double* __restrict__ acc, uint32_t* __restrict__ touched,
                                int* __restrict__ touched_n) {
  const int j = threadIdx.x;
  if (j >= k) return;
  const int64_t r = knn_rows[j];
  if (r < 0) return;
  const uint32_t d = (uint32_t)(r - row_base);
  if (row_filter && ((int64_t)d >= filter_rows || !row_filter[d])) return;   // bool.filter drops the neighbour
  const float c = __fmul_rn(w_knn, knn_scores[j]);
  const double old = atomicAdd(acc + d, (double)c);
  if (old == 0.0) touched[atomicAdd(touched_n, 1)] = d;
}

// per-warp top lists over the touched rows; clears the accumulator behind itself
template <int M>
__global__ void __launch_bounds__(RASS_WARPS_PER_CTA * 32) fuse_collect_kernel(
    double* __restrict__ acc, const uint32_t* __restrict__ touched, const int* __restrict__ touched_n,
    double* __restrict__ xkey, uint32_t* __restrict__ xrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * RASS_WARPS_PER_CTA + warp;
  const int W = gridDim.x * RASS_WARPS_PER_CTA;
  const int n = *touched_n;
  WarpTop<double, M> top;
  top.init();
  for (int base = gw * 32; base < n; base += W * 32) {
    const int i = base + lane;
    double key = neg_inf<double>();
    uint32_t d = 0xffffffffu;
    if (i < n) {
      d = touched[i];
      key = (double)(float)acc[d];   // Lucene casts the summed clause scores to float
      acc[d] = 0.0;
    }
    unsigned hit = __ballot_sync(0xffffffffu, i < n && entry_better<double>(key, d, top.thr_key, top.thr_row));
    while (hit) {
      const int src = __ffs(hit) - 1;
      hit &= hit - 1;
      top.insert(shfl_t(key, src), __shfl_sync(0xffffffffu, d, src));
    }
  }
#pragma unroll
  for (int s = 0; s < M; ++s) {
    xkey[(size_t)gw * (32 * M) + s * 32 + lane] = top.key[s];
    xrow[(size_t)gw * (32 * M) + s * 32 + lane] = top.row[s];
  }
}

__global__ void reset_counter_kernel(int* p) { *p = 0; }

int launch_select_raw(rass_engine* h, size_t entries, int k, int q, int64_t* out_rows, float* out_scores,
                      cudaStream_t st);

extern "C" int rass_search_hybrid(rass_engine* h, const float* q_host, int B, const int32_t* qterm_indptr,
                                  const int32_t* qterms, float w_text, float w_knn, int k, int64_t* out_rows,
                                  float* out_scores, rass_stats* stats) {
  if (!h) return RASS_E_INVALID;
  cudaSetDevice(h->device);
  if (B < 1 || !out_rows || !out_scores) return rass_fail(h, RASS_E_INVALID, "bad arguments");
  if (k < 1 || k > RASS_MAX_K) return rass_fail(h, RASS_E_INVALID, "k must be in [1, %d], got %d", RASS_MAX_K, k);
  if (!q_host && !qterm_indptr) return rass_fail(h, RASS_E_INVALID, "neither a vector nor a text clause");
  Bm25State& b = h->bm25;
  if (qterm_indptr && (!b.acc || !qterms)) return rass_fail(h, RASS_E_INVALID, "text clause without rass_bm25_build");
  cudaStream_t st = eng_stream(h);
  int rc;
  const size_t n_out = (size_t)B * k;
  if ((rc = ensure_out_workspace(h, 2 * n_out))) return rc;
  // accumulator must cover the vector rows too
  const int64_t need = std::max<int64_t>(std::max<int64_t>(b.N, h->n_rows), 1);
  if (need > b.acc_rows) {
    CUDA_TRY(h, cudaStreamSynchronize(st));
    cudaFree(b.acc); b.acc = nullptr;
    cudaFree(b.touched); b.touched = nullptr;
    CUDA_TRY(h, cudaMalloc(&b.acc, (size_t)need * sizeof(double)));
    CUDA_TRY(h, cudaMemset(b.acc, 0, (size_t)need * sizeof(double)));
    CUDA_TRY(h, cudaMalloc(&b.touched, (size_t)need * sizeof(uint32_t)));
    if (!b.touched_n) {
      CUDA_TRY(h, cudaMalloc(&b.touched_n, sizeof(int)));
      CUDA_TRY(h, cudaMemset(b.touched_n, 0, sizeof(int)));
    }
    b.acc_rows = need;
    b.touched_cap = need;
  }
  rass_stats s;
  memset(&s, 0, sizeof(s));
  s.n_queries = B;
  // 1. the knn clause: exact top-k per query (results stay on the device in the second half of the staging)
  int64_t* knn_rows = h->out_rows + n_out;
  float* knn_scores = h->out_scores + n_out;
  const bool have_vec = q_host != nullptr && h->n_rows > 0;
  if (have_vec) {
    float* q_dev = nullptr;
    if ((rc = stage_queries(h, q_host, B, &q_dev))) return rc;
    if ((rc = search_core(h, q_dev, B, k, knn_rows, knn_scores, nullptr, &s))) return rc;
  }
  // 2. per query: postings walk, knn contributions, top-k over the touched rows
  const int M = k <= 32 ? 1 : 4;
  const int grid_c = h->num_sms;
  const size_t entries = (size_t)grid_c * RASS_WARPS_PER_CTA * 32 * M;
  if ((rc = ensure_xlist_workspace(h, entries))) return rc;
  cudaEvent_t e0 = h->ev[0], e1 = h->ev[3];
  CUDA_TRY(h, cudaEventRecord(e0, st));
  for (int q = 0; q < B; ++q) {
    if (qterm_indptr) {
      TermArgs ta;
      memset(&ta, 0, sizeof(ta));
      const float bo = w_text;
      for (int32_t j = qterm_indptr[q]; j < qterm_indptr[q + 1]; ++j) {
        const int32_t t = qterms[j];
        if (t < 0 || t >= b.V) continue;
        const int64_t lo = b.indptr_host[(size_t)t], len = b.indptr_host[(size_t)t + 1] - lo;
        if (len == 0) continue;
        if (ta.n_terms == BM25_MAX_TERMS) return rass_fail(h, RASS_E_INVALID, "more than %d query terms", BM25_MAX_TERMS);
        ta.lo[ta.n_terms] = lo;
        ta.cum[ta.n_terms + 1] = ta.cum[ta.n_terms] + len;
        volatile float w = bo * b.idf_host[(size_t)t];
        ta.w[ta.n_terms] = w;
        ++ta.n_terms;
      }
      const int64_t total = ta.cum[ta.n_terms];
      if (total > 0) {
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 16);
        bm25_accumulate_kernel<<<blocks, 256, 0, st>>>(ta, b.doc, b.tf, b.norm, b.inv_dev, h->sb, h->n_rows,
                                                       h->row_filter, h->row_filter_rows, b.acc, b.touched,
                                                       b.touched_n);
        CUDA_TRY(h, cudaGetLastError());
        s.launches++;
        s.bytes_streamed += total * 6;
      }
    }
    if (have_vec) {
      fuse_knn_kernel<<<1, RASS_MAX_K, 0, st>>>(knn_rows + (size_t)q * k, knn_scores + (size_t)q * k, k, h->row_base,
                                                w_knn, h->row_filter, h->row_filter_rows, b.acc, b.touched,
                                                b.touched_n);
      CUDA_TRY(h, cudaGetLastError());
      s.launches++;
    }
    if (M == 1) fuse_collect_kernel<1><<<grid_c, RASS_WARPS_PER_CTA * 32, 0, st>>>(b.acc, b.touched, b.touched_n, h->xlist_key, h->xlist_row);
    else fuse_collect_kernel<4><<<grid_c, RASS_WARPS_PER_CTA * 32, 0, st>>>(b.acc, b.touched, b.touched_n, h->xlist_key, h->xlist_row);
    CUDA_TRY(h, cudaGetLastError());
    if ((rc = launch_select_raw(h, entries, k, q, h->out_rows, h->out_scores, st))) return rc;
    reset_counter_kernel<<<1, 1, 0, st>>>(b.touched_n);
    CUDA_TRY(h, cudaGetLastError());
    s.launches += 3;
  }
  CUDA_TRY(h, cudaEventRecord(e1, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->out_rows_host, h->out_rows, n_out * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->out_scores_host, h->out_scores, n_out * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  memcpy(out_rows, h->out_rows_host, n_out * 8);
  memcpy(out_scores, h->out_scores_host, n_out * 4);
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
  s.finish_ms += ms;
  s.total_ms += ms;
  if (stats) *stats = s;
  return RASS_OK;
}
