// quintnet_b200 — fused AdamW on flat buffers (K8).
// One launch updates a whole bucket: fp32 m/v (+ optional fp32 master), bf16 or fp32 parameters and
// gradients, decoupled weight decay, bias correction, and an optional gradient scale read from a DEVICE
// scalar (the clip coefficient) so clipping needs no host round-trip.  Pure streaming: 16-byte vectors.
#include "common.cuh"

namespace qn {

template <typename PT, typename GT>
__global__ void __launch_bounds__(256)
adamw_kernel(PT* __restrict__ p, const GT* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             float* __restrict__ master, const float* __restrict__ grad_scale, long long n, float lr, float beta1,
             float beta2, float eps, float wd, float inv_bc1, float inv_sqrt_bc2) {
  const float gs = grad_scale != nullptr ? *grad_scale : 1.0f;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float gf[4], pf[4];
    if constexpr (sizeof(GT) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(g + i * 4);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      gf[0] = a.x; gf[1] = a.y; gf[2] = b.x; gf[3] = b.y;
    } else {
      const float4 u = *reinterpret_cast<const float4*>(g + i * 4);
      gf[0] = u.x; gf[1] = u.y; gf[2] = u.z; gf[3] = u.w;
    }
    if (master != nullptr) {
      const float4 u = *reinterpret_cast<const float4*>(master + i * 4);
      pf[0] = u.x; pf[1] = u.y; pf[2] = u.z; pf[3] = u.w;
    } else if constexpr (sizeof(PT) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(p + i * 4);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      pf[0] = a.x; pf[1] = a.y; pf[2] = b.x; pf[3] = b.y;
    } else {
      const float4 u = *reinterpret_cast<const float4*>(p + i * 4);
      pf[0] = u.x; pf[1] = u.y; pf[2] = u.z; pf[3] = u.w;
    }
    float4 mm = *reinterpret_cast<const float4*>(m + i * 4);
    float4 vv = *reinterpret_cast<const float4*>(v + i * 4);
    float* mp = reinterpret_cast<float*>(&mm);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gf[e] * gs;
      mp[e] = beta1 * mp[e] + (1.0f - beta1) * gr;
      vp[e] = beta2 * vp[e] + (1.0f - beta2) * gr * gr;
      const float denom = sqrtf(vp[e]) * inv_sqrt_bc2 + eps;
      pf[e] = pf[e] * (1.0f - lr * wd) - lr * (mp[e] * inv_bc1) / denom;
    }
    *reinterpret_cast<float4*>(m + i * 4) = mm;
    *reinterpret_cast<float4*>(v + i * 4) = vv;
    if (master != nullptr) *reinterpret_cast<float4*>(master + i * 4) = make_float4(pf[0], pf[1], pf[2], pf[3]);
    if constexpr (sizeof(PT) == 2) {
      uint2 o;
      o.x = pack_bf16(pf[0], pf[1]); o.y = pack_bf16(pf[2], pf[3]);
      *reinterpret_cast<uint2*>(p + i * 4) = o;
    } else {
      *reinterpret_cast<float4*>(p + i * 4) = make_float4(pf[0], pf[1], pf[2], pf[3]);
    }
  }
}

}  // namespace qn

// pdtype/gdtype: 0 = fp32, 1 = bf16.  n must be a multiple of 4 (flat buckets are padded to 128).
extern "C" int qn_adamw(void* p, int pdtype, const void* g, int gdtype, float* m, float* v, float* master,
                        const float* grad_scale, long long n, float lr, float beta1, float beta2, float eps,
                        float wd, float bc1, float bc2, cudaStream_t s) {
  using namespace qn;
  if (n % 4) return 940;
  long long grid = (n / 4 + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  const float inv_bc1 = 1.0f / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
#define QN_ADAMW(PT, GT)                                                                                        \
  adamw_kernel<PT, GT><<<(int)grid, 256, 0, s>>>((PT*)p, (const GT*)g, m, v, master, grad_scale, n, lr, beta1,  \
                                                 beta2, eps, wd, inv_bc1, inv_sqrt_bc2)
  if (pdtype == 1 && gdtype == 1) QN_ADAMW(__nv_bfloat16, __nv_bfloat16);
  else if (pdtype == 1 && gdtype == 0) QN_ADAMW(__nv_bfloat16, float);
  else if (pdtype == 0 && gdtype == 1) QN_ADAMW(float, __nv_bfloat16);
  else QN_ADAMW(float, float);
#undef QN_ADAMW
  return (int)cudaGetLastError();
}
