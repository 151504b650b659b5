// quintnet_b200 — fused elementwise kernels (bf16 I/O, fp32 math, 16-byte vectors):
//   * dropout(x)·1/(1-p) + residual   forward / backward (mask regenerated from a counter hash)
//   * GELU (exact erf) forward / backward (+ optional bias add)
//   * column sum (bias gradient)
//   * RoPE forward / backward (K7)
#include "common.cuh"

namespace qn {

constexpr int EW_THREADS = 256;

QN_DEVICE void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
QN_DEVICE void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// y = dropout(x) * scale + residual   (residual may be null)
__global__ void __launch_bounds__(EW_THREADS)
dropout_add_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                       __nv_bfloat16* __restrict__ y, long long n8, uint32_t threshold, float scale, uint64_t seed,
                       uint64_t offset) {
  const uint32_t dkey = drop_key32(seed, offset);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], r[8];
    load8(x + i * 8, f);
    const uint32_t m = keep_mask8_fast(dkey, i, threshold);
    if (res != nullptr) load8(res + i * 8, r);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = ((m >> e) & 1 ? f[e] * scale : 0.f) + (res != nullptr ? r[e] : 0.f);
    store8(y + i * 8, f);
  }
}
// dx = dy * mask * scale
__global__ void __launch_bounds__(EW_THREADS)
dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, long long n8,
                   uint32_t threshold, float scale, uint64_t seed, uint64_t offset) {
  const uint32_t dkey = drop_key32(seed, offset);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    load8(dy + i * 8, f);
    const uint32_t m = keep_mask8_fast(dkey, i, threshold);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (m >> e) & 1 ? f[e] * scale : 0.f;
    store8(dx + i * 8, f);
  }
}

__global__ void __launch_bounds__(EW_THREADS)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    load8(x + i * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = gelu_erf(f[e]);
    store8(y + i * 8, f);
  }
}
// dx = dy * gelu'(z)
__global__ void __launch_bounds__(EW_THREADS)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                __nv_bfloat16* __restrict__ dx, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], zz[8];
    load8(dy + i * 8, g);
    load8(z + i * 8, zz);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] *= gelu_erf_grad(zz[e]);
    store8(dx + i * 8, g);
  }
}

// out[c] = sum_r x[r, c]  (bias gradient).  grid.x tiles columns (64 per CTA: 8 vector-columns of 8 bf16,
// 32 row-lanes), grid.y splits rows; partial sums are combined with fp32 atomics into `acc`
// (zeroed by the caller), then converted.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ acc, int rows, int cols, long long ld) {
  __shared__ float red[32][65];
  const int cv = threadIdx.x & 7;          // vector column within the 64-wide tile
  const int rl = threadIdx.x >> 3;         // row lane 0..31
  const int c0 = blockIdx.x * 64 + cv * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < cols) {
    // four independent 16-byte loads in flight per thread: the sweep is latency-bound otherwise
    const int step = gridDim.y * 32;
    int r = blockIdx.y * 32 + rl;
    for (; r + 3 * step < rows; r += 4 * step) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_nc_v4(x + (size_t)(r + k * step) * ld + c0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
        s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
      }
    }
    for (; r < rows; r += step) {
      float f[8];
      load8(x + (size_t)r * ld + c0, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cv * 8 + e] = s[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(acc + c, t);
  }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}

// RoPE (rotate-half): x viewed as [n_rows, T, d] with arbitrary row/seq strides (elements); cos/sin [T, d/2] fp32.
// sign = +1 forward, -1 backward (the rotation is orthogonal: the gradient is the inverse rotation)
__global__ void __launch_bounds__(EW_THREADS)
rope_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ cosp,
            const float* __restrict__ sinp, int B, int H, int T, int d, long long sb, long long sh, long long st,
            float sign) {
  const int half = d / 2;
  const long long total = (long long)B * H * T * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = i % half;
    long long r = i / half;
    const int t = r % T; r /= T;
    const int h = r % H;
    const int b = r / H;
    const __nv_bfloat16* xp = x + b * sb + h * sh + t * st;
    const float x1 = __bfloat162float(xp[j]), x2 = __bfloat162float(xp[j + half]);
    const float c = cosp[t * half + j], s = sign * sinp[t * half + j];
    __nv_bfloat16* yp = y + (((long long)b * H + h) * T + t) * d;       // output contiguous [B,H,T,d]
    yp[j] = __float2bfloat16(x1 * c - x2 * s);
    yp[j + half] = __float2bfloat16(x2 * c + x1 * s);
  }
}

static int ew_grid(long long n) {
  long long g = (n + EW_THREADS - 1) / EW_THREADS;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

}  // namespace qn

using namespace qn;

extern "C" int qn_dropout_add_fwd(const void* x, const void* res, void* y, long long n, float p,
                                  unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  if (n % 8) return 930;
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  dropout_add_fwd_kernel<<<ew_grid(n / 8), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)res,
                                                              (__nv_bfloat16*)y, n / 8, thr, 1.0f / (1.0f - thr / 65536.0f), seed, offset);
  return (int)cudaGetLastError();
}
extern "C" int qn_dropout_bwd(const void* dy, void* dx, long long n, float p, unsigned long long seed,
                              unsigned long long offset, cudaStream_t s) {
  if (n % 8) return 930;
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  dropout_bwd_kernel<<<ew_grid(n / 8), EW_THREADS, 0, s>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, n / 8, thr,
                                                          1.0f / (1.0f - thr / 65536.0f), seed, offset);
  return (int)cudaGetLastError();
}
extern "C" int qn_gelu_fwd(const void* x, void* y, long long n, cudaStream_t s) {
  if (n % 8) return 930;
  gelu_fwd_kernel<<<ew_grid(n / 8), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8);
  return (int)cudaGetLastError();
}
extern "C" int qn_gelu_bwd(const void* dy, const void* z, void* dx, long long n, cudaStream_t s) {
  if (n % 8) return 930;
  gelu_bwd_kernel<<<ew_grid(n / 8), EW_THREADS, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)z,
                                                       (__nv_bfloat16*)dx, n / 8);
  return (int)cudaGetLastError();
}
// acc: fp32 [cols] workspace (zeroed here); out: bf16 [cols]
extern "C" int qn_colsum(const void* x, float* acc, void* out, int rows, int cols, long long ld, cudaStream_t s) {
  if ((cols % 8) || (ld % 8)) return 930;
  cudaMemsetAsync(acc, 0, sizeof(float) * cols, s);
  int gy = (rows + 31) / 32;
  const int gx = (cols + 63) / 64;
  const int cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  colsum_kernel<<<dim3(gx, gy), 256, 0, s>>>((const __nv_bfloat16*)x, acc, rows, cols, ld);
  f32_to_bf16_kernel<<<(cols + 255) / 256, 256, 0, s>>>(acc, (__nv_bfloat16*)out, cols);
  return (int)cudaGetLastError();
}
// acc32[cols] += column sums of x (fp32 accumulation straight into a main-grad buffer: no memset, no conversion)
extern "C" int qn_colsum_accum_f32(const void* x, float* acc32, int rows, int cols, long long ld, cudaStream_t s) {
  if ((cols % 8) || (ld % 8)) return 930;
  int gy = (rows + 31) / 32;
  const int gx = (cols + 63) / 64;
  const int cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  colsum_kernel<<<dim3(gx, gy), 256, 0, s>>>((const __nv_bfloat16*)x, acc32, rows, cols, ld);
  return (int)cudaGetLastError();
}
extern "C" int qn_rope(const void* x, void* y, const float* cosp, const float* sinp, int B, int H, int T, int d,
                       long long sb, long long sh, long long st, int backward, cudaStream_t s) {
  if (d % 2) return 930;
  const long long total = (long long)B * H * T * (d / 2);
  rope_kernel<<<ew_grid(total), EW_THREADS, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, cosp, sinp, B, H, T, d,
                                                   sb, sh, st, backward ? -1.0f : 1.0f);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Row-wise FP8 (e4m3) quantisation for the FP8 GEMM: one warp per row, amax over the row, q = x * 448 / amax,
// scale_out[row] = amax / 448 (the dequantisation factor the GEMM epilogue multiplies back in).
// ---------------------------------------------------------------------------------------------------------
#include <cuda_fp8.h>
namespace qn {
__global__ void __launch_bounds__(256)
quantize_rowwise_e4m3_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, float* __restrict__ scale_out,
                             int rows, int cols, long long ldx, long long ldq) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  const int nvec = cols >> 3;
  float amax = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    load8(xr + v * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float inv = amax > 0.f ? 448.0f / amax : 1.0f;
  if (lane == 0) scale_out[row] = amax > 0.f ? amax / 448.0f : 1.0f;
  uint8_t* qr = q + (size_t)row * ldq;
  for (int v = lane; v < nvec; v += 32) {                 // second sweep hits L1/L2
    float f[8];
    load8(xr + v * 8, f);
    uint2 o;
    const uint32_t p0 = __nv_cvt_float2_to_fp8x2(make_float2(f[0] * inv, f[1] * inv), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p1 = __nv_cvt_float2_to_fp8x2(make_float2(f[2] * inv, f[3] * inv), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p2 = __nv_cvt_float2_to_fp8x2(make_float2(f[4] * inv, f[5] * inv), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p3 = __nv_cvt_float2_to_fp8x2(make_float2(f[6] * inv, f[7] * inv), __NV_SATFINITE, __NV_E4M3);
    o.x = p0 | (p1 << 16);
    o.y = p2 | (p3 << 16);
    *reinterpret_cast<uint2*>(qr + v * 8) = o;
  }
}
}  // namespace qn

extern "C" int qn_quantize_rowwise_e4m3(const void* x, void* q, float* scale_out, int rows, int cols, long long ldx,
                                        long long ldq, cudaStream_t s) {
  if ((cols % 8) || (ldx % 8) || (ldq % 8)) return 940;
  qn::quantize_rowwise_e4m3_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)x, (uint8_t*)q, scale_out, rows,
                                                                 cols, ldx, ldq);
  return (int)cudaGetLastError();
}
