// Vector-store fill and query preparation kernels.
//
// store_convert replaces what OpenSearch does with `_source["embedding"]` on a bulk "index" action
// (reference app/main.py:1256-1269): it lays one appended row down as the resident fp32 row, its bf16
// shadow (the array the scan streams), its fp64-accumulated norm and the per-row scan scale/offset.
// query_prep is the device half of the query normalisation at app/main.py:1536-1537.
#include "common.cuh"

// one warp per row
__global__ void __launch_bounds__(256) store_convert_kernel(const float* __restrict__ src, int64_t src_stride, int dim,
                                                            int dim_pad, int metric, int bf16_only, int write32,
                                                            float* __restrict__ x32, __nv_bfloat16* __restrict__ x16,
                                                            double* __restrict__ norm64, float* __restrict__ sa,
                                                            float* __restrict__ sb, DevScalars* scal,
                                                            int64_t first_row, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t row = first_row + r;
  const float* s = src + r * src_stride;
  double n2 = 0.0, e2 = 0.0;
  for (int j = lane * 4; j < dim_pad; j += 128) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (j + e < dim) ? s[j + e] : 0.f;
    __nv_bfloat16 b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      b[e] = __float2bfloat16_rn(v[e]);
      float back = __bfloat162float(b[e]);
      if (bf16_only) {
        n2 = fma((double)back, (double)back, n2);
      } else {
        n2 = fma((double)v[e], (double)v[e], n2);
        double d = (double)v[e] - (double)back;
        e2 = fma(d, d, e2);
      }
    }
    uint2 packed;
    packed.x = (uint32_t)__bfloat16_as_ushort(b[0]) | ((uint32_t)__bfloat16_as_ushort(b[1]) << 16);
    packed.y = (uint32_t)__bfloat16_as_ushort(b[2]) | ((uint32_t)__bfloat16_as_ushort(b[3]) << 16);
    *reinterpret_cast<uint2*>(x16 + (size_t)row * dim_pad + j) = packed;
    if (write32) *reinterpret_cast<float4*>(x32 + (size_t)row * dim_pad + j) = make_float4(v[0], v[1], v[2], v[3]);
  }
  n2 = warp_sum(n2);
  e2 = warp_sum(e2);
  if (lane == 0) {
    double nrm = sqrt(n2);
    norm64[row] = nrm;
    if (metric == RASS_METRIC_COSINE) {
      sa[row] = nrm > 0.0 ? (float)(1.0 / nrm) : 0.f;
      sb[row] = 0.f;
    } else {
      sa[row] = 1.f;
      sb[row] = (float)(-0.5 * n2);
    }
    // round the ratios up so the bound stays a bound after the float conversion
    float rho = nrm > 0.0 ? __double2float_ru(sqrt(e2) / nrm) : 0.f;
    atomicMax(reinterpret_cast<int*>(&scal->rho_x), __float_as_int(rho));
    atomicMax(reinterpret_cast<int*>(&scal->max_xnorm), __float_as_int(__double2float_ru(nrm)));
  }
}

int launch_store_convert(rass_engine* h, const float* src_dev, int64_t src_stride, int64_t first_row, int64_t n,
                         cudaStream_t st) {
  if (n <= 0) return RASS_OK;
  h->sb_filtered_dirty = true;
  const bool bf16_only = (h->flags & RASS_BF16_ONLY) != 0;
  const bool in_place = !bf16_only && src_dev == h->x32 + (size_t)first_row * h->dim_pad;
  const int warps = 8;
  int64_t blocks = (n + warps - 1) / warps;
  store_convert_kernel<<<(unsigned)blocks, warps * 32, 0, st>>>(src_dev, src_stride, h->dim, h->dim_pad, h->metric,
                                                                bf16_only ? 1 : 0, (!bf16_only && !in_place) ? 1 : 0,
                                                                h->x32, h->x16, h->norm64, h->sa, h->sb, h->scal,
                                                                first_row, n);
  CUDA_TRY(h, cudaGetLastError());
  return RASS_OK;
}

// one warp per query slot; slots >= B (up to the 256-multiple the tcgen05 passes read) are zeroed
__global__ void __launch_bounds__(128) query_prep_kernel(const float* __restrict__ q, int dim, int dim_pad, int metric,
                                                         int B, int B_pad, float* __restrict__ q_raw,
                                                         float* __restrict__ q_hat, __nv_bfloat16* __restrict__ q16,
                                                         double* __restrict__ q_norm, float* __restrict__ q_rho) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B_pad) return;
  if (b >= B) {
    for (int j = lane; j < dim_pad; j += 32) q16[(size_t)b * dim_pad + j] = __float2bfloat16_rn(0.f);
    return;
  }
  const float* s = q + (size_t)b * dim;
  double n2 = 0.0;
  for (int j = lane; j < dim; j += 32) n2 = fma((double)s[j], (double)s[j], n2);
  n2 = warp_sum(n2);
  const double nrm = sqrt(n2);
  const double inv = (metric == RASS_METRIC_COSINE) ? (nrm > 0.0 ? 1.0 / nrm : 0.0) : 1.0;
  double h2 = 0.0, e2 = 0.0;
  for (int j = lane; j < dim_pad; j += 32) {
    float v = j < dim ? s[j] : 0.f;
    float hat = (float)((double)v * inv);
    __nv_bfloat16 hb = __float2bfloat16_rn(hat);
    double d = (double)hat - (double)__bfloat162float(hb);
    h2 = fma((double)hat, (double)hat, h2);
    e2 = fma(d, d, e2);
    q_raw[(size_t)b * dim_pad + j] = v;
    q_hat[(size_t)b * dim_pad + j] = hat;
    q16[(size_t)b * dim_pad + j] = hb;
  }
  h2 = warp_sum(h2);
  e2 = warp_sum(e2);
  if (lane == 0) {
    q_norm[b] = nrm;
    q_rho[b] = h2 > 0.0 ? __double2float_ru(sqrt(e2 / h2)) : 0.f;
  }
}

int launch_query_prep(rass_engine* h, const float* q_dev, int B, cudaStream_t st) {
  const int B_pad = (B + RASS_QPAD - 1) / RASS_QPAD * RASS_QPAD;
  const int warps = 4;
  query_prep_kernel<<<(B_pad + warps - 1) / warps, warps * 32, 0, st>>>(q_dev, h->dim, h->dim_pad, h->metric, B, B_pad,
                                                                        h->q_raw, h->q_hat, h->q16, h->q_norm,
                                                                        h->q_rho);
  CUDA_TRY(h, cudaGetLastError());
  return RASS_OK;
}
