// quintnet_b200 — LayerNorm forward / backward (K6), optionally fused with the residual add.
//
// Memory-bound: one warp per row, 16-byte vector loads, row kept in registers, warp-shuffle reductions only
// (no shared memory, no __syncthreads in the forward).  Backward accumulates dgamma/dbeta per warp in
// registers over a grid-stride loop of rows, reduces across the CTA through shared memory, and a second
// tiny kernel folds the per-CTA partials.
#include "common.cuh"

#include <cstdlib>

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


// ---------------------------------------------------------------------------------------------------------
// Fused backward for bf16 (the training path): dx (+ dropout'(dx)), dgamma and dbeta in ONE pass over dy / x / dadd.
// Row blocks of 8 rows are streamed into shared memory by 1-D bulk async copies (rows are contiguous, so a block
// is one copy per tensor) through a 2-stage mbarrier ring, which keeps ~100 KB per SM in flight without holding
// it in registers.  Phase 1: one warp per row computes the two row means.  Phase 2: each thread owns 8 columns
// (one 16-byte vector) for half of the rows, writes dx with fully coalesced stores and accumulates dgamma / dbeta
// for its columns in 16 registers over every row the CTA ever sees; one vector `red` per thread at the end.
// ---------------------------------------------------------------------------------------------------------
constexpr int LNF_R = 8, LNF_STAGES = 2, LNF_THREADS = 256;

template <int CV, bool HAS_ADD>
__global__ void __launch_bounds__(LNF_THREADS)
ln_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
                    const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dadd,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_drop, float* __restrict__ acc,
                    int rows, int cols, const LnDropout dr, const int want_colsum, float* __restrict__ acc_g,
                    float* __restrict__ acc_b) {
  // dgamma / dbeta accumulators: acc / acc + cols by default, or caller-provided fp32 buffers (the parameters' main
  // grads: the kernel then accumulates the LayerNorm weight gradients in place, nothing to finalize or add later)
  if (acc_g == nullptr) { acc_g = acc; acc_b = acc + cols; }
  extern __shared__ __align__(128) uint8_t lnf_smem[];
  __shared__ uint64_t full[LNF_STAGES];
  __shared__ float s_stat[LNF_STAGES][LNF_R][4];          // mean, rstd, c1, c2 per row
  constexpr int NT = HAS_ADD ? 3 : 2;
  const int row_bytes = cols * 2;
  const int tens_bytes = LNF_R * row_bytes;
  const int stage_bytes = NT * tens_bytes;
  uint8_t* s_gamma = lnf_smem + LNF_STAGES * stage_bytes;
  const int nvec = cols >> 3;
  const int nblk = (rows + LNF_R - 1) / LNF_R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid & 127, h = tid >> 7;
  const uint32_t dkey = drop_key32(dr.seed, dr.offset);

  for (int i = tid; i < nvec; i += LNF_THREADS)
    reinterpret_cast<uint4*>(s_gamma)[i] = reinterpret_cast<const uint4*>(gamma)[i];
  if (tid == 0) {
    for (int i = 0; i < LNF_STAGES; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int blk, int st) {
    const int r0 = blk * LNF_R;
    const uint32_t bytes = static_cast<uint32_t>(min(LNF_R, rows - r0)) * row_bytes;
    uint8_t* sb = lnf_smem + st * stage_bytes;
    mbar_arrive_expect_tx(&full[st], bytes * NT);
    bulk_load_1d(sb, dy + static_cast<size_t>(r0) * cols, bytes, &full[st]);
    bulk_load_1d(sb + tens_bytes, x + static_cast<size_t>(r0) * cols, bytes, &full[st]);
    if (HAS_ADD) bulk_load_1d(sb + 2 * tens_bytes, dadd + static_cast<size_t>(r0) * cols, bytes, &full[st]);
  };
  if (tid == 0) {
    for (int st = 0; st < LNF_STAGES; ++st) {
      const int blk = blockIdx.x + st * gridDim.x;
      if (blk < nblk) issue(blk, st);
    }
  }
  // ad: column sums of the gradient that leaves through the branch input (dx_drop, or dx without dropout) = the
  // bias gradient of the Linear layer that produced the branch — saves that layer a separate column-sum pass
  float g2[CV][8], ag[CV][8], ab[CV][8], ad[CV][8];
#pragma unroll
  for (int j = 0; j < CV; ++j) {
    const int vi = c + 128 * j;
#pragma unroll
    for (int e = 0; e < 8; ++e) { ag[j][e] = 0.f; ab[j][e] = 0.f; g2[j][e] = 0.f; ad[j][e] = 0.f; }
    if (vi < nvec) Vec<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(s_gamma) + vi * 8, g2[j]);
  }

  int it = 0;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++it) {
    const int st = it % LNF_STAGES;
    const uint32_t ph = (it / LNF_STAGES) & 1;
    const int r0 = blk * LNF_R;
    const int nr = min(LNF_R, rows - r0);
    float mean = 0.f, rstd = 0.f;
    if (warp < nr) { mean = mean_in[r0 + warp]; rstd = rstd_in[r0 + warp]; }   // latency hides behind the wait
    mbar_wait(&full[st], ph);
    const uint8_t* sb = lnf_smem + st * stage_bytes;
    // ---- phase 1: warp -> row: c1 = mean(dy*g), c2 = mean(dy*g*xhat)
    if (warp < nr) {
      const __nv_bfloat16* sdy = reinterpret_cast<const __nv_bfloat16*>(sb + warp * row_bytes);
      const __nv_bfloat16* sx = reinterpret_cast<const __nv_bfloat16*>(sb + tens_bytes + warp * row_bytes);
      float c1 = 0.f, c2 = 0.f;
      for (int vi = lane; vi < nvec; vi += 32) {
        float d[8], xv[8], g[8];
        Vec<__nv_bfloat16>::load(sdy + vi * 8, d);
        Vec<__nv_bfloat16>::load(sx + vi * 8, xv);
        Vec<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(s_gamma) + vi * 8, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dyg = d[e] * g[e];
          c1 += dyg;
          c2 = fmaf(dyg, (xv[e] - mean) * rstd, c2);
        }
      }
      c1 = warp_sum(c1) / cols;
      c2 = warp_sum(c2) / cols;
      if (lane == 0) { s_stat[st][warp][0] = mean; s_stat[st][warp][1] = rstd; s_stat[st][warp][2] = c1; s_stat[st][warp][3] = c2; }
    }
    __syncthreads();
    // ---- phase 2: thread -> 8 columns x 4 rows
#pragma unroll 1
    for (int rr = 0; rr < LNF_R / 2; ++rr) {
      const int r = h * (LNF_R / 2) + rr;
      if (r >= nr) break;
      const float rmean = s_stat[st][r][0], rrstd = s_stat[st][r][1], c1 = s_stat[st][r][2], c2 = s_stat[st][r][3];
      const size_t grow = static_cast<size_t>(r0 + r);
#pragma unroll
      for (int j = 0; j < CV; ++j) {
        const int vi = c + 128 * j;
        if (vi < nvec) {
          float d[8], xv[8], o[8];
          Vec<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(sb + r * row_bytes) + vi * 8, d);
          Vec<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(sb + tens_bytes + r * row_bytes) + vi * 8, xv);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = (xv[e] - rmean) * rrstd;
            ag[j][e] = fmaf(d[e], xh, ag[j][e]);
            ab[j][e] += d[e];
            o[e] = rrstd * (d[e] * g2[j][e] - c1 - xh * c2);
          }
          if (HAS_ADD) {
            float a[8];
            Vec<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(sb + 2 * tens_bytes + r * row_bytes) + vi * 8, a);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += a[e];
          }
          Vec<__nv_bfloat16>::store(dx + grow * cols + vi * 8, o);
          if (dx_drop != nullptr) {
            if (dr.threshold != 0) {
              const uint32_t m = keep_mask8_fast(dkey, grow * nvec + vi, dr.threshold);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = ((m >> e) & 1) ? o[e] * dr.scale : 0.f;
            }
            Vec<__nv_bfloat16>::store(dx_drop + grow * cols + vi * 8, o);
          }
          if (want_colsum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) ad[j][e] += __bfloat162float(__float2bfloat16(o[e]));   // sum what was stored
          }
        }
      }
    }
    __syncthreads();                                       // every thread is done with this stage
    if (tid == 0) {
      const int nb = blk + LNF_STAGES * gridDim.x;
      if (nb < nblk) issue(nb, st);
    }
  }
  // ---- dgamma / dbeta: fold the two row halves through shared memory, one vector reduction per thread
  __syncthreads();
  float* red = reinterpret_cast<float*>(lnf_smem);          // [128 * CV][24]; the ring is idle now
  if (h == 1) {
#pragma unroll
    for (int j = 0; j < CV; ++j) {
      const int vi = c + 128 * j;
      if (vi < nvec) {                                     // the scratch is only nvec * 96 bytes deep
#pragma unroll
        for (int e = 0; e < 8; ++e) { red[vi * 24 + e] = ag[j][e]; red[vi * 24 + 8 + e] = ab[j][e]; red[vi * 24 + 16 + e] = ad[j][e]; }
      }
    }
  }
  __syncthreads();
  if (h == 0) {
#pragma unroll
    for (int j = 0; j < CV; ++j) {
      const int vi = c + 128 * j;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { ag[j][e] += red[vi * 24 + e]; ab[j][e] += red[vi * 24 + 8 + e]; ad[j][e] += red[vi * 24 + 16 + e]; }
        if (want_colsum) {
#pragma unroll
          for (int q = 0; q < 2; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(acc + 2 * cols + vi * 8 + q * 4), "f"(ad[j][q * 4 + 0]),
                         "f"(ad[j][q * 4 + 1]), "f"(ad[j][q * 4 + 2]), "f"(ad[j][q * 4 + 3]) : "memory");
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(acc_g + vi * 8 + q * 4), "f"(ag[j][q * 4 + 0]),
                       "f"(ag[j][q * 4 + 1]), "f"(ag[j][q * 4 + 2]), "f"(ag[j][q * 4 + 3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(acc_b + vi * 8 + q * 4), "f"(ab[j][q * 4 + 0]),
                       "f"(ab[j][q * 4 + 1]), "f"(ab[j][q * 4 + 2]), "f"(ab[j][q * 4 + 3]) : "memory");
        }
      }
    }
  }
}

template <int CV, bool HAS_ADD>
static int launch_ln_bwd_fused(const void* dy, const void* x, const void* g, const float* mean, const float* rstd,
                               const void* dadd, void* dx, void* dx_drop, float* acc, int rows, int cols,
                               const LnDropout dr, cudaStream_t s, int want_colsum = 0, float* acc_g = nullptr,
                               float* acc_b = nullptr) {
  const int nt = HAS_ADD ? 3 : 2;
  const int smem = LNF_STAGES * nt * LNF_R * cols * 2 + cols * 2;
  auto k = ln_bwd_fused_kernel<CV, HAS_ADD>;
  static bool attr_done = false;
  if (!attr_done) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
  const int per_sm = max(1, min(4, (220 * 1024) / (smem + 1024)));
  const int nblk = (rows + LNF_R - 1) / LNF_R;
  const int grid = min(nblk, 148 * per_sm);
  using B = __nv_bfloat16;
  k<<<grid, LNF_THREADS, smem, s>>>((const B*)dy, (const B*)x, (const B*)g, mean, rstd, (const B*)dadd, (B*)dx,
                                   (B*)dx_drop, acc, rows, cols, dr, want_colsum, acc_g, acc_b);
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
__global__ void ln_finalize_kernel(const float* __restrict__ acc, T* __restrict__ dgamma, T* __restrict__ dbeta, int cols,
                                   T* __restrict__ colsum = nullptr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) {
    if (dgamma != nullptr) { dgamma[c] = static_cast<T>(acc[c]); dbeta[c] = static_cast<T>(acc[cols + c]); }
    if (colsum != nullptr) colsum[c] = static_cast<T>(acc[2 * cols + c]);
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
                  int cols, void* dx_drop, const LnDropout dr, cudaStream_t s, void* colsum = nullptr,
                  float* main_g = nullptr, float* main_b = nullptr) {
  constexpr int VN = Vec<T>::N;
  if (cols % VN || cols > 32 * LN_MAXV * VN) return 920;
  (void)nparts;
  if (sizeof(T) == 2 && cols <= 2048 && rows >= 64 && getenv("QN_LN_BWD_SPLIT") == nullptr) {
    // fused single-pass kernel (bf16): fp32 accumulators [2 * cols] in the workspace, then bf16 conversion
    const int wc = colsum != nullptr ? 1 : 0;
    const bool in_place = main_g != nullptr && main_b != nullptr;      // dgamma / dbeta go straight into fp32 main grads
    if (!in_place || wc) cudaMemsetAsync(workspace, 0, (2ull + wc) * cols * sizeof(float), s);
    int rc;
    if (cols <= 1024) rc = dadd ? launch_ln_bwd_fused<1, true>(dy, x, g, mean, rstd, dadd, dx, dx_drop, workspace, rows, cols, dr, s, wc, main_g, main_b)
                                : launch_ln_bwd_fused<1, false>(dy, x, g, mean, rstd, dadd, dx, dx_drop, workspace, rows, cols, dr, s, wc, main_g, main_b);
    else rc = dadd ? launch_ln_bwd_fused<2, true>(dy, x, g, mean, rstd, dadd, dx, dx_drop, workspace, rows, cols, dr, s, wc, main_g, main_b)
                   : launch_ln_bwd_fused<2, false>(dy, x, g, mean, rstd, dadd, dx, dx_drop, workspace, rows, cols, dr, s, wc, main_g, main_b);
    if (rc) return rc;
    if (in_place && !wc) return 0;
    ln_finalize_kernel<T><<<(cols + 255) / 256, 256, 0, s>>>(workspace, in_place ? nullptr : (T*)dgamma, (T*)dbeta, cols, (T*)colsum);
    return static_cast<int>(cudaGetLastError());
  }
  if (colsum != nullptr || main_g != nullptr) return 921;   // only the fused single-pass kernel has these outputs
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
//   colsum (optional, bf16 [cols]): column sums of the gradient w.r.t. the branch input x (dx, or dres when p == 0) — the
//   bias gradient of the Linear that produced x; needs the fused kernel (cols <= 2048, rows >= 64), else error 921.
extern "C" int qn_dropout_add_layernorm_bwd(const void* dy, const void* h, const void* gamma, const float* mean,
                                            const float* rstd, const void* dsum, void* dres, void* dx, void* dgamma,
                                            void* dbeta, float* workspace, int nparts, int rows, int cols, float p,
                                            unsigned long long seed, unsigned long long offset, void* colsum,
                                            float* main_g, float* main_b, cudaStream_t s) {
  //   main_g / main_b (optional, fp32 [cols]): the LayerNorm weight / bias MAIN-GRAD buffers — dgamma / dbeta are
  //   accumulated into them in place (dgamma / dbeta outputs are then not written)
  qn::LnDropout dr{0u, 1.0f, seed, offset};
  if (p > 0.f) { dr.threshold = (uint32_t)(p * 65536.0f); dr.scale = 1.0f / (1.0f - dr.threshold / 65536.0f); }
  return qn::ln_bwd<__nv_bfloat16>(dy, h, gamma, mean, rstd, dsum, dres, dgamma, dbeta, workspace, nparts, rows, cols,
                                   dx, dr, s, colsum, main_g, main_b);
}
