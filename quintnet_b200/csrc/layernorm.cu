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

struct LnDropout {
  uint32_t threshold;   // 16-bit keep threshold (p * 65536); 0 = no dropout
  float scale;          // 1 / (1 - threshold / 65536)
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
  const uint32_t dkey = drop_key32(dr.seed, dr.offset);
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
          const uint32_t m = keep_mask8_fast(dkey, static_cast<uint64_t>(row) * nvec + vi, dr.threshold);
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

// ---------------------------------------------------------------------------------------------- backward
// Split in two high-occupancy kernels (the single-pass version needed 201 registers -> 2 CTAs/SM):
//   (A) dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)) [+ dadd]   (+ optional dropout'(dx) output)
//   (B) dgamma = sum_rows dy * xhat, dbeta = sum_rows dy   — column reduction with coalesced row sweeps
template <typename T, int MAXV, bool HAS_ADD>
__global__ void __launch_bounds__(LN_WARPS * 32, 4)
layernorm_bwd_dx_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gamma,
                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                        const T* __restrict__ dadd, T* __restrict__ dx, int rows, int cols, T* __restrict__ dx_drop,
                        const LnDropout dr) {
  constexpr int VN = Vec<T>::N;
  const uint32_t dkey = drop_key32(dr.seed, dr.offset);
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * LN_WARPS;
  const int nvec = cols / VN;
  for (int row = warp_global; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[MAXV][VN], dyg[MAXV][VN];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float g[VN];
        Vec<T>::load(x + static_cast<size_t>(row) * cols + vi * VN, xh[i]);
        Vec<T>::load(dy + static_cast<size_t>(row) * cols + vi * VN, dyg[i]);
        Vec<T>::load(gamma + vi * VN, g);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          xh[i][e] = (xh[i][e] - mean) * rstd;
          dyg[i][e] *= g[e];
          c1 += dyg[i][e];
          c2 = fmaf(dyg[i][e], xh[i][e], c2);
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
        for (int e = 0; e < VN; ++e) o[e] = rstd * (dyg[i][e] - c1 - xh[i][e] * c2);
        if (HAS_ADD) {
          float a[VN];
          Vec<T>::load(dadd + static_cast<size_t>(row) * cols + vi * VN, a);
#pragma unroll
          for (int e = 0; e < VN; ++e) o[e] += a[e];
        }
        Vec<T>::store(dx + static_cast<size_t>(row) * cols + vi * VN, o);
        if (dx_drop != nullptr) {                            // gradient of the dropped branch input
          if (VN == 8 && dr.threshold != 0) {
            const uint32_t m = keep_mask8_fast(dkey, static_cast<uint64_t>(row) * nvec + vi, dr.threshold);
#pragma unroll
            for (int e = 0; e < VN; ++e) o[e] = ((m >> e) & 1) ? o[e] * dr.scale : 0.f;
          }
          Vec<T>::store(dx_drop + static_cast<size_t>(row) * cols + vi * VN, o);
        }
      }
    }
  }
}

// block = 256 threads = (cols/VN vector-columns: 8) x (32 row lanes); grid.x tiles columns, grid.y splits rows
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_dgdb_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, float* __restrict__ acc, int rows, int cols) {
  constexpr int VN = Vec<T>::N;
  constexpr int TW = 8 * VN;                    // columns per CTA
  __shared__ float sg[32][TW + 1], sb[32][TW + 1];
  const int cv = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * TW + cv * VN;
  float g[VN], b[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { g[e] = 0.f; b[e] = 0.f; }
  if (c0 < cols) {
    for (int r = blockIdx.y * 32 + rl; r < rows; r += gridDim.y * 32) {
      float xv[VN], dv[VN];
      Vec<T>::load(x + static_cast<size_t>(r) * cols + c0, xv);
      Vec<T>::load(dy + static_cast<size_t>(r) * cols + c0, dv);
      const float mean = mean_in[r], rstd = rstd_in[r];
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        g[e] = fmaf(dv[e], (xv[e] - mean) * rstd, g[e]);
        b[e] += dv[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) { sg[rl][cv * VN + e] = g[e]; sb[rl][cv * VN + e] = b[e]; }
  __syncthreads();
  if (threadIdx.x < TW) {
    float tg = 0.f, tb = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { tg += sg[r][threadIdx.x]; tb += sb[r][threadIdx.x]; }
    const int c = blockIdx.x * TW + threadIdx.x;
    if (c < cols) { atomicAdd(acc + c, tg); atomicAdd(acc + cols + c, tb); }
  }
}

template <typename T>
__global__ void ln_finalize_kernel(const float* __restrict__ acc, T* __restrict__ dgamma, T* __restrict__ dbeta, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) { dgamma[c] = static_cast<T>(acc[c]); dbeta[c] = static_cast<T>(acc[cols + c]); }
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
  (void)nparts;
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (grid > 148 * 16) grid = 148 * 16;
  const int mv = (cols / VN + 31) / 32;
#define QN_LN_BWD(MV, ADD)                                                                                            \
  layernorm_bwd_dx_kernel<T, MV, ADD><<<grid, LN_WARPS * 32, 0, s>>>((const T*)dy, (const T*)x, (const T*)g, mean,    \
                                                                     rstd, (const T*)dadd, (T*)dx, rows, cols,       \
                                                                     (T*)dx_drop, dr)
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
  // dgamma / dbeta: fp32 accumulators [2 * cols] in the workspace
  cudaMemsetAsync(workspace, 0, 2ull * cols * sizeof(float), s);
  constexpr int TW = 8 * VN;
  const int gx = (cols + TW - 1) / TW;
  int gy = (rows + 31) / 32;
  const int cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  layernorm_bwd_dgdb_kernel<T><<<dim3(gx, gy), 256, 0, s>>>((const T*)dy, (const T*)x, mean, rstd, workspace, rows, cols);
  ln_finalize_kernel<T><<<(cols + 255) / 256, 256, 0, s>>>(workspace, (T*)dgamma, (T*)dbeta, cols);
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
  if (p > 0.f) { dr.threshold = (uint32_t)(p * 65536.0f); dr.scale = 1.0f / (1.0f - dr.threshold / 65536.0f); }
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
  if (p > 0.f) { dr.threshold = (uint32_t)(p * 65536.0f); dr.scale = 1.0f / (1.0f - dr.threshold / 65536.0f); }
  return qn::ln_bwd<__nv_bfloat16>(dy, h, gamma, mean, rstd, dsum, dres, dgamma, dbeta, workspace, nparts, rows, cols,
                                   dx, dr, s);
}
