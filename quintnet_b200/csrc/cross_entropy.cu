// quintnet_b200 — fused softmax cross-entropy forward+backward on bf16 logits (part of the fused
// lm_head + CE op).  One CTA per token row:
//   pass 1: online max / sum-exp over the vocabulary (16-byte loads, fp32 math)
//   pass 2: overwrite the row IN PLACE with  (softmax - onehot)  in bf16  (zero for ignored rows)
// and accumulate  sum(loss)  /  n_valid  with one atomic per row.  The fp32 [N, V] logits tensor of the
// reference (3.3 GB at 32x512x50257) never exists; the bf16 logits buffer becomes the gradient buffer.
#include "common.cuh"

namespace qn {

constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                  float* __restrict__ loss_sum, float* __restrict__ n_valid, int V, long long ld, long long ignore_index,
                  int write_grad) {
  __shared__ float s_max[CE_THREADS / 32], s_sum[CE_THREADS / 32];
  const long long row = blockIdx.x;
  __nv_bfloat16* x = logits + row * ld;
  const long long label = labels[row];
  const bool valid = (label != ignore_index);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nv = V / 8;  // full 16-byte vectors; tail handled scalar
  if (!valid) {
    if (write_grad) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < nv; i += CE_THREADS) *reinterpret_cast<uint4*>(x + i * 8) = z;
      for (int c = nv * 8 + tid; c < V; c += CE_THREADS) x[c] = __float2bfloat16(0.f);
    }
    return;
  }
  // ---- pass 1: online softmax statistics
  float mx = -INFINITY, sm = 0.f;
  for (int i = tid; i < nv; i += CE_THREADS) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    const float lm = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)), fmaxf(fmaxf(c.x, c.y), fmaxf(d.x, d.y)));
    if (lm > mx) { sm *= __expf(mx - lm); mx = lm; }
    sm += __expf(a.x - mx) + __expf(a.y - mx) + __expf(b.x - mx) + __expf(b.y - mx) + __expf(c.x - mx) +
          __expf(c.y - mx) + __expf(d.x - mx) + __expf(d.y - mx);
  }
  for (int c = nv * 8 + tid; c < V; c += CE_THREADS) {
    const float f = __bfloat162float(x[c]);
    if (f > mx) { sm *= __expf(mx - f); mx = f; }
    sm += __expf(f - mx);
  }
  // warp then CTA combine of (max, sum)
  // (threads / warps that saw no element carry max = -inf, sum = 0: guard the inf - inf)
  const float wm = warp_max(mx);
  sm = warp_sum(mx == -INFINITY ? 0.f : sm * __expf(mx - wm));
  if (lane == 0) { s_max[w] = wm; s_sum[w] = sm; }
  __syncthreads();
  float gm = s_max[0];
#pragma unroll
  for (int i = 1; i < CE_THREADS / 32; ++i) gm = fmaxf(gm, s_max[i]);
  float gs = 0.f;
#pragma unroll
  for (int i = 0; i < CE_THREADS / 32; ++i) gs += (s_max[i] == -INFINITY) ? 0.f : s_sum[i] * __expf(s_max[i] - gm);
  const float lse = gm + __logf(gs);
  if (tid == 0) {
    const float xl = __bfloat162float(x[label]);
    atomicAdd(loss_sum, lse - xl);
    atomicAdd(n_valid, 1.0f);
  }
  if (!write_grad) return;
  __syncthreads();  // x[label] read above must precede the overwrite
  // ---- pass 2: gradient in place
  const float inv = 1.0f / gs;
  for (int i = tid; i < nv; i += CE_THREADS) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
    const long long base = (long long)i * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f[e] = __expf(f[e] - gm) * inv;
      if (base + e == label) f[e] -= 1.0f;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(x + i * 8) = o;
  }
  for (int c = nv * 8 + tid; c < V; c += CE_THREADS) {
    float f = __expf(__bfloat162float(x[c]) - gm) * inv;
    if (c == label) f -= 1.0f;
    x[c] = __float2bfloat16(f);
  }
}

// y[i] = x[i] * (*scale)   (bf16, in place allowed): applies the upstream d(loss) scalar to dlogits-derived grads
__global__ void __launch_bounds__(256)
scale_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                  long long n8) {
  const float s = *scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    uint4 o;
    o.x = pack_bf16(a.x * s, a.y * s); o.y = pack_bf16(b.x * s, b.y * s);
    o.z = pack_bf16(c.x * s, c.y * s); o.w = pack_bf16(d.x * s, d.y * s);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

}  // namespace qn

// logits: bf16 [rows, ld] (ld % 8 == 0, 16-byte aligned rows); labels int64 [rows];
// loss_sum / n_valid: fp32 device scalars, accumulated into (caller zeroes them)
extern "C" int qn_cross_entropy(void* logits, const long long* labels, float* loss_sum, float* n_valid, int rows,
                                int V, long long ld, long long ignore_index, int write_grad, cudaStream_t s) {
  if (ld % 8) return 950;
  if (rows <= 0) return 0;
  qn::ce_fwd_bwd_kernel<<<rows, qn::CE_THREADS, 0, s>>>((__nv_bfloat16*)logits, labels, loss_sum, n_valid, V, ld,
                                                        ignore_index, write_grad);
  return (int)cudaGetLastError();
}
extern "C" int qn_scale_bf16(const void* x, void* y, const float* scale, long long n, cudaStream_t s) {
  if (n % 8) return 950;
  long long grid = (n / 8 + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  qn::scale_bf16_kernel<<<(int)grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, scale, n / 8);
  return (int)cudaGetLastError();
}
