// quintnet_b200 — LayerNorm forward / backward (K6), optionally fused with the residual add.
//
// Memory-bound: one warp per row, 16-byte vector loads, row kept in registers, warp-shuffle reductions only
// (no shared memory, no __syncthreads in the forward).  Backward accumulates dgamma/dbeta per warp in
// registers over a grid-stride loop of rows, reduces across the CTA through shared memory, and a second
// tiny kernel folds the per-CTA partials.
#include "common.cuh"

namespace qn {

template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static QN_DEVICE void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static QN_DEVICE void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  static QN_DEVICE void load(const float* p, float (&f)[4]) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
  }
  static QN_DEVICE void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

// keep-mask for the 8 consecutive elements of 16-byte vector `i8` (identical in forward and backward)
QN_DEVICE uint32_t ln_keep_mask8(const Philox& rng, uint64_t offset, uint64_t i8, uint32_t threshold) {
  const uint4 a = rng(offset + 2 * i8), b = rng(offset + 2 * i8 + 1);
  uint32_t m = 0;
  m |= (a.x >= threshold) << 0; m |= (a.y >= threshold) << 1; m |= (a.z >= threshold) << 2; m |= (a.w >= threshold) << 3;
  m |= (b.x >= threshold) << 4; m |= (b.y >= threshold) << 5; m |= (b.z >= threshold) << 6; m |= (b.w >= threshold) << 7;
  return m;
}
struct LnDropout {
  uint32_t threshold;   // 0 = no dropout
  float scale;          // 1 / (1 - p)
  uint64_t seed, offset;
};

constexpr int LN_WARPS = 4;
constexpr int LN_MAXV = 8;   // max vectors per lane kept in registers: cols <= 32*8*VecN (2048 bf16 / 1024 fp32)

// y = LN(x [+ residual]) ; optionally writes sum = x + residual ; saves mean / rstd (fp32) for backward
template <typename T, int MAXV, bool HAS_RES>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ gamma,
                     const T* __restrict__ beta, T* __restrict__ y, T* __restrict__ sum_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int cols, float eps,
                     const LnDropout dr) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * LN_WARPS;
  const int nvec = cols / VN;
  const Philox rng(dr.seed);
  for (int row = warp_global; row < rows; row += nwarps) {
    const T* xr = x + static_cast<size_t>(row) * cols;
    float v[MAXV][VN];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        Vec<T>::load(xr + vi * VN, v[i]);
        if (HAS_RES && VN == 8 && dr.threshold != 0) {       // y = LN(dropout(x)·s + residual)
          const uint32_t m = ln_keep_mask8(rng, dr.offset, static_cast<uint64_t>(row) * nvec + vi, dr.threshold);
#pragma unroll
          for (int e = 0; e < VN; ++e) v[i][e] = ((m >> e) & 1) ? v[i][e] * dr.scale : 0.f;
        }
        if (HAS_RES) {
          float r[VN];
          Vec<T>::load(res + static_cast<size_t>(row) * cols + vi * VN, r);
#pragma unroll
          for (int e = 0; e < VN; ++e) v[i][e] += r[e];
          if (sum_out != nullptr) Vec<T>::store(sum_out + static_cast<size_t>(row) * cols + vi * VN, v[i]);
        }
#pragma unroll
        for (int e = 0; e < VN; ++e) s += v[i][e];
      }
    }
    const float mean = warp_sum(s) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int e = 0; e < VN; ++e) { const float d = v[i][e] - mean; sq += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0 && mean_out != nullptr) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float g[VN], b[VN], o[VN];
        Vec<T>::load(gamma + vi * VN, g);
        Vec<T>::load(beta + vi * VN, b);
#pragma unroll
        for (int e = 0; e < VN; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
        Vec<T>::store(y + static_cast<size_t>(row) * cols + vi * VN, o);
      }
    }
  }
}

// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)) [+ dres_in];  partial dgamma/dbeta per CTA
template <typename T, int MAXV, bool HAS_ADD>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const T* __restrict__ dadd, T* __restrict__ dx, float* __restrict__ dgamma_part,
                     float* __restrict__ dbeta_part, int rows, int cols, T* __restrict__ dx_drop, const LnDropout dr) {
  constexpr int VN = Vec<T>::N;
  const Philox rng(dr.seed);
  extern __shared__ float red[];  // [LN_WARPS][cols] x 2
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * LN_WARPS + w;
  const int nwarps = gridDim.x * LN_WARPS;
  const int nvec = cols / VN;
  float dg[MAXV][VN], db[MAXV][VN], g[MAXV][VN];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
#pragma unroll
    for (int e = 0; e < VN; ++e) { dg[i][e] = 0.f; db[i][e] = 0.f; g[i][e] = 0.f; }
    if (vi < nvec) Vec<T>::load(gamma + vi * VN, g[i]);
  }
  for (int row = warp_global; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[MAXV][VN], dyv[MAXV][VN];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        Vec<T>::load(x + static_cast<size_t>(row) * cols + vi * VN, xh[i]);
        Vec<T>::load(dy + static_cast<size_t>(row) * cols + vi * VN, dyv[i]);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          xh[i][e] = (xh[i][e] - mean) * rstd;
          const float dyg = dyv[i][e] * g[i][e];
          c1 += dyg;
          c2 += dyg * xh[i][e];
          dg[i][e] += dyv[i][e] * xh[i][e];
          db[i][e] += dyv[i][e];
        }
      }
    }
    c1 = warp_sum(c1) / cols;
    c2 = warp_sum(c2) / cols;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float o[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) o[e] = rstd * (dyv[i][e] * g[i][e] - c1 - xh[i][e] * c2);
        if (HAS_ADD) {
          float a[VN];
          Vec<T>::load(dadd + static_cast<size_t>(row) * cols + vi * VN, a);
#pragma unroll
          for (int e = 0; e < VN; ++e) o[e] += a[e];
        }
        Vec<T>::store(dx + static_cast<size_t>(row) * cols + vi * VN, o);
        if (dx_drop != nullptr) {                            // gradient of the dropped branch input
          if (VN == 8 && dr.threshold != 0) {
            const uint32_t m = ln_keep_mask8(rng, dr.offset, static_cast<uint64_t>(row) * nvec + vi, dr.threshold);
#pragma unroll
            for (int e = 0; e < VN; ++e) o[e] = ((m >> e) & 1) ? o[e] * dr.scale : 0.f;
          }
          Vec<T>::store(dx_drop + static_cast<size_t>(row) * cols + vi * VN, o);
        }
      }
    }
  }
  float* rg = red;
  float* rb = red + LN_WARPS * cols;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        rg[w * cols + vi * VN + e] = dg[i][e];
        rb[w * cols + vi * VN + e] = db[i][e];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int ww = 0; ww < LN_WARPS; ++ww) { a += rg[ww * cols + c]; b += rb[ww * cols + c]; }
    dgamma_part[static_cast<size_t>(blockIdx.x) * cols + c] = a;
    dbeta_part[static_cast<size_t>(blockIdx.x) * cols + c] = b;
  }
}

// fold the per-CTA partials: block = 32 columns x 8 partial-lanes
template <typename T>
__global__ void __launch_bounds__(256)
ln_reduce_partials_kernel(const float* __restrict__ pg, const float* __restrict__ pb,
                          T* __restrict__ dgamma, T* __restrict__ dbeta, int nparts, int cols) {
  __shared__ float sg[8][33], sb[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f, b = 0.f;
  if (c < cols) {
    for (int i = py; i < nparts; i += 8) {
      a += pg[static_cast<size_t>(i) * cols + c];
      b += pb[static_cast<size_t>(i) * cols + c];
    }
  }
  sg[py][cx] = a; sb[py][cx] = b;
  __syncthreads();
  if (py == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a += sg[i][cx]; b += sb[i][cx]; }
    dgamma[c] = static_cast<T>(a);
    dbeta[c] = static_cast<T>(b);
  }
}

template <typename T>
static int ln_fwd(const void* x, const void* res, const void* g, const void* b, void* y, void* sum_out, float* mean,
                  float* rstd, int rows, int cols, float eps, const LnDropout dr, cudaStream_t s) {
  constexpr int VN = Vec<T>::N;
  if (cols % VN || cols > 32 * LN_MAXV * VN) return 920;
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (grid > 148 * 16) grid = 148 * 16;
  const int mv = (cols / VN + 31) / 32;
#define QN_LN_FWD(MV)                                                                                                  \
  do {                                                                                                                 \
    if (res != nullptr)                                                                                                \
      layernorm_fwd_kernel<T, MV, true><<<grid, LN_WARPS * 32, 0, s>>>((const T*)x, (const T*)res, (const T*)g,        \
                                                                       (const T*)b, (T*)y, (T*)sum_out, mean, rstd,   \
                                                                       rows, cols, eps, dr);                          \
    else                                                                                                               \
      layernorm_fwd_kernel<T, MV, false><<<grid, LN_WARPS * 32, 0, s>>>((const T*)x, nullptr, (const T*)g,             \
                                                                        (const T*)b, (T*)y, nullptr, mean, rstd, rows, \
                                                                        cols, eps, dr);                                \
  } while (0)
  if (mv <= 1) QN_LN_FWD(1); else if (mv <= 2) QN_LN_FWD(2); else if (mv <= 4) QN_LN_FWD(4); else QN_LN_FWD(8);
#undef QN_LN_FWD
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
static int ln_bwd(const void* dy, const void* x, const void* g, const float* mean, const float* rstd,
                  const void* dadd, void* dx, void* dgamma, void* dbeta, float* workspace, int nparts, int rows,
                  int cols, void* dx_drop, const LnDropout dr, cudaStream_t s) {
  constexpr int VN = Vec<T>::N;
  if (cols % VN || cols > 32 * LN_MAXV * VN) return 920;
  const size_t smem = 2ull * LN_WARPS * cols * sizeof(float);
  float* pg = workspace;
  float* pb = workspace + static_cast<size_t>(nparts) * cols;
  const int mv = (cols / VN + 31) / 32;
#define QN_LN_BWD(MV, ADD)                                                                                            \
  do {                                                                                                                 \
    auto k = layernorm_bwd_kernel<T, MV, ADD>;                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    k<<<nparts, LN_WARPS * 32, smem, s>>>((const T*)dy, (const T*)x, (const T*)g, mean, rstd, (const T*)dadd, (T*)dx, \
                                          pg, pb, rows, cols, (T*)dx_drop, dr);                                       \
  } while (0)
#define QN_LN_BWD2(ADD)                                                                                               \
  do {                                                                                                                 \
    if (mv <= 1) QN_LN_BWD(1, ADD); else if (mv <= 2) QN_LN_BWD(2, ADD); else if (mv <= 4) QN_LN_BWD(4, ADD);         \
    else QN_LN_BWD(8, ADD);                                                                                           \
  } while (0)
  if (dadd != nullptr) QN_LN_BWD2(true); else QN_LN_BWD2(false);
#undef QN_LN_BWD2
#undef QN_LN_BWD
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  ln_reduce_partials_kernel<T><<<(cols + 31) / 32, 256, 0, s>>>(pg, pb, (T*)dgamma, (T*)dbeta, nparts, cols);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace qn

// dtype: 0 = fp32, 1 = bf16 (x, residual, gamma, beta, y share it)
extern "C" int qn_layernorm_fwd(const void* x, const void* res, const void* gamma, const void* beta, void* y,
                                void* sum_out, float* mean, float* rstd, int rows, int cols, float eps, int dtype,
                                cudaStream_t s) {
  const qn::LnDropout nd{0u, 1.0f, 0ull, 0ull};
  if (dtype == 1) return qn::ln_fwd<__nv_bfloat16>(x, res, gamma, beta, y, sum_out, mean, rstd, rows, cols, eps, nd, s);
  return qn::ln_fwd<float>(x, res, gamma, beta, y, sum_out, mean, rstd, rows, cols, eps, nd, s);
}

// y = LN(h),  h = dropout(x, p)·1/(1-p) + res  (h written to sum_out); bf16 only
extern "C" int qn_dropout_add_layernorm_fwd(const void* x, const void* res, const void* gamma, const void* beta,
                                            void* y, void* sum_out, float* mean, float* rstd, int rows, int cols,
                                            float eps, float p, unsigned long long seed, unsigned long long offset,
                                            cudaStream_t s) {
  qn::LnDropout dr{0u, 1.0f, seed, offset};
  if (p > 0.f) { dr.threshold = (uint32_t)fminf(4294967295.0f, p * 4294967296.0f); dr.scale = 1.0f / (1.0f - p); }
  return qn::ln_fwd<__nv_bfloat16>(x, res, gamma, beta, y, sum_out, mean, rstd, rows, cols, eps, dr, s);
}

// workspace: fp32 [2 * nparts * cols]; nparts = number of CTAs (<= 148*4 is plenty)
extern "C" int qn_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                const float* rstd, const void* dadd, void* dx, void* dgamma, void* dbeta,
                                float* workspace, int nparts, int rows, int cols, int dtype, cudaStream_t s) {
  const qn::LnDropout nd{0u, 1.0f, 0ull, 0ull};
  if (dtype == 1)
    return qn::ln_bwd<__nv_bfloat16>(dy, x, gamma, mean, rstd, dadd, dx, dgamma, dbeta, workspace, nparts, rows, cols,
                                     nullptr, nd, s);
  return qn::ln_bwd<float>(dy, x, gamma, mean, rstd, dadd, dx, dgamma, dbeta, workspace, nparts, rows, cols, nullptr, nd, s);
}

// backward of qn_dropout_add_layernorm_fwd:  dres = LN'(dy) + dsum ;  dx = dropout'(dres)
extern "C" int qn_dropout_add_layernorm_bwd(const void* dy, const void* h, const void* gamma, const float* mean,
                                            const float* rstd, const void* dsum, void* dres, void* dx, void* dgamma,
                                            void* dbeta, float* workspace, int nparts, int rows, int cols, float p,
                                            unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  qn::LnDropout dr{0u, 1.0f, seed, offset};
  if (p > 0.f) { dr.threshold = (uint32_t)fminf(4294967295.0f, p * 4294967296.0f); dr.scale = 1.0f / (1.0f - p); }
  return qn::ln_bwd<__nv_bfloat16>(dy, h, gamma, mean, rstd, dsum, dres, dgamma, dbeta, workspace, nparts, rows, cols,
                                   dx, dr, s);
}
