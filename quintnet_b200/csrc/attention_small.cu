// quintnet_b200 — whole-head attention for SHORT sequences (ViT-MNIST: T = 50 tokens, head_dim 16): one CTA per
// (batch, head) keeps Q, K, V and the full [T x T] score matrix in shared memory, so forward and backward are one
// launch each with no HBM round trip for the scores and no head split / transpose / concat copies (inputs and
// outputs are addressed through strides; the natural layout is the [B, T, H*d] output of the Q / K / V projections).
//
// This is the latency-bound regime: 8 images x 4 heads x 50 x 50 scores is ~10^5 MACs per head — far
// below one tcgen05 tile — so the design goal is launch count and bytes, not tensor-core throughput; the big-model
// path (T >= 128, head_dim 64) is the tcgen05 flash kernel in attention.cu.
//
// Replaces F.scaled_dot_product_attention on the ViT path (reference utils/model.py:45-110 computes the same
// softmax(QKᵀ/√d)V with torch ops).  fp32 math in shared memory, bf16 I/O.
#include "common.cuh"

namespace qn {
namespace sattn {

constexpr int THREADS = 256;

struct Params {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;   // element (b, h, t, :) at b*sb + h*sh + t*st
  long long q_sb, q_sh, q_st, k_sb, k_sh, k_st, v_sb, v_sh, v_st;
  __nv_bfloat16* o; long long o_sb, o_sh, o_st;                             // fwd: output; bwd: the forward output
  const __nv_bfloat16* d_o; long long do_sb, do_sh, do_st;                  // bwd: upstream gradient (layout of o)
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv;                  // bwd: gradients, laid out like q / k / v
  float* lse;                                                               // [B, H, T] log-sum-exp of the scaled scores
  int B, H, T, D;
  float scale;
  int causal;
};

QN_DEVICE void load_tile(float* dst, const __nv_bfloat16* src, long long st, int T, int D, int ld) {
  // [T, D] bf16 rows (row stride st) -> fp32 smem rows of pitch ld; 8 elements (16 B) per thread-iteration
  const int vec_per_row = D / 8;
  for (int i = threadIdx.x; i < T * vec_per_row; i += THREADS) {
    const int t = i / vec_per_row, c = (i % vec_per_row) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(src + t * st + c);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
    float* p = dst + t * ld + c;
    p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y; p[4] = cc.x; p[5] = cc.y; p[6] = d.x; p[7] = d.y;
  }
}
QN_DEVICE void store_tile(__nv_bfloat16* dst, long long st, const float* src, int T, int D, int ld) {
  const int vec_per_row = D / 8;
  for (int i = threadIdx.x; i < T * vec_per_row; i += THREADS) {
    const int t = i / vec_per_row, c = (i % vec_per_row) * 8;
    const float* p = src + t * ld + c;
    uint4 u;
    u.x = pack_bf16(p[0], p[1]); u.y = pack_bf16(p[2], p[3]); u.z = pack_bf16(p[4], p[5]); u.w = pack_bf16(p[6], p[7]);
    *reinterpret_cast<uint4*>(dst + t * st + c) = u;
  }
}

// S[i][j] = scale * q_i . k_j (masked -> -inf), then row softmax in place; returns with P in sS and lse in sL
QN_DEVICE void scores_softmax(const float* sQ, const float* sK, float* sS, float* sL, int T, int D, int ld, int lds,
                              float scale, int causal) {
  for (int e = threadIdx.x; e < T * T; e += THREADS) {
    const int i = e / T, j = e % T;
    float acc = 0.f;
    const float* qi = sQ + i * ld; const float* kj = sK + j * ld;
#pragma unroll 8
    for (int c = 0; c < D; ++c) acc = fmaf(qi[c], kj[c], acc);
    sS[i * lds + j] = (causal && j > i) ? -INFINITY : acc * scale;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < T; i += THREADS / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sS[i * lds + j]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = __expf(sS[i * lds + j] - mx);
      sS[i * lds + j] = e;
      sm += e;
    }
    sm = warp_sum(sm);
    const float inv = 1.0f / sm;
    for (int j = lane; j < T; j += 32) sS[i * lds + j] *= inv;
    if (lane == 0) sL[i] = mx + __logf(sm);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS)
small_attn_fwd_kernel(const Params p) {
  extern __shared__ float sm[];
  const int T = p.T, D = p.D, ld = D + 1, lds = T + 1;
  float* sQ = sm; float* sK = sQ + T * ld; float* sV = sK + T * ld; float* sS = sV + T * ld; float* sL = sS + T * lds;
  const int h = blockIdx.x, b = blockIdx.y;
  load_tile(sQ, p.q + b * p.q_sb + h * p.q_sh, p.q_st, T, D, ld);
  load_tile(sK, p.k + b * p.k_sb + h * p.k_sh, p.k_st, T, D, ld);
  load_tile(sV, p.v + b * p.v_sb + h * p.v_sh, p.v_st, T, D, ld);
  __syncthreads();
  scores_softmax(sQ, sK, sS, sL, T, D, ld, lds, p.scale, p.causal);
  // O = P V, written over sQ (no longer needed)
  for (int e = threadIdx.x; e < T * D; e += THREADS) {
    const int i = e / D, c = e % D;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(sS[i * lds + j], sV[j * ld + c], acc);
    sQ[i * ld + c] = acc;
  }
  __syncthreads();
  store_tile(p.o + b * p.o_sb + h * p.o_sh, p.o_st, sQ, T, D, ld);
  if (p.lse != nullptr)
    for (int i = threadIdx.x; i < T; i += THREADS) p.lse[(static_cast<long long>(b) * p.H + h) * T + i] = sL[i];
}

__global__ void __launch_bounds__(THREADS)
small_attn_bwd_kernel(const Params p) {
  extern __shared__ float sm[];
  const int T = p.T, D = p.D, ld = D + 1, lds = T + 1;
  float* sQ = sm; float* sK = sQ + T * ld; float* sV = sK + T * ld; float* sdO = sV + T * ld;
  float* sP = sdO + T * ld; float* sdS = sP + T * lds; float* sL = sdS + T * lds;
  const int h = blockIdx.x, b = blockIdx.y;
  load_tile(sQ, p.q + b * p.q_sb + h * p.q_sh, p.q_st, T, D, ld);
  load_tile(sK, p.k + b * p.k_sb + h * p.k_sh, p.k_st, T, D, ld);
  load_tile(sV, p.v + b * p.v_sb + h * p.v_sh, p.v_st, T, D, ld);
  load_tile(sdO, p.d_o + b * p.do_sb + h * p.do_sh, p.do_st, T, D, ld);
  __syncthreads();
  scores_softmax(sQ, sK, sP, sL, T, D, ld, lds, p.scale, p.causal);       // recompute P (exactly the forward's)
  // dP[i][j] = dO_i . V_j
  for (int e = threadIdx.x; e < T * T; e += THREADS) {
    const int i = e / T, j = e % T;
    float acc = 0.f;
    const float* a = sdO + i * ld; const float* vj = sV + j * ld;
#pragma unroll 8
    for (int c = 0; c < D; ++c) acc = fmaf(a[c], vj[c], acc);
    sdS[i * lds + j] = acc;
  }
  __syncthreads();
  // delta_i = sum_j P_ij dP_ij ;  dS = P * (dP - delta) * scale
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < T; i += THREADS / 32) {
    float d = 0.f;
    for (int j = lane; j < T; j += 32) d += sP[i * lds + j] * sdS[i * lds + j];
    d = warp_sum(d);
    for (int j = lane; j < T; j += 32) sdS[i * lds + j] = sP[i * lds + j] * (sdS[i * lds + j] - d) * p.scale;
  }
  __syncthreads();
  // dV_j = sum_i P_ij dO_i ;  dK_j = sum_i dS_ij Q_i ;  dQ_i = sum_j dS_ij K_j  — one output element per thread-iteration,
  // written straight to global memory (2400 elements per head at T = 50, d = 16: not worth a staging pass)
  for (int e = threadIdx.x; e < T * D; e += THREADS) {
    const int r = e / D, c = e % D;
    float dv = 0.f, dk = 0.f, dq = 0.f;
    for (int i = 0; i < T; ++i) {
      dv = fmaf(sP[i * lds + r], sdO[i * ld + c], dv);
      dk = fmaf(sdS[i * lds + r], sQ[i * ld + c], dk);
      dq = fmaf(sdS[r * lds + i], sK[i * ld + c], dq);
    }
    p.dv[b * p.v_sb + h * p.v_sh + r * p.v_st + c] = __float2bfloat16(dv);
    p.dk[b * p.k_sb + h * p.k_sh + r * p.k_st + c] = __float2bfloat16(dk);
    p.dq[b * p.q_sb + h * p.q_sh + r * p.q_st + c] = __float2bfloat16(dq);
  }
}

}  // namespace sattn
}  // namespace qn

// q, k, v: bf16, element (b, h, t, c) at  b*sb + h*sh + t*st + c  (c contiguous, D % 8 == 0, 16-byte aligned rows).
// o / d_o likewise.  lse: fp32 [B, H, T] (may be null in forward when no backward will follow).
static size_t sattn_smem(int T, int D, bool bwd) {
  const size_t ld = D + 1, lds = T + 1;
  return sizeof(float) * (bwd ? (4 * T * ld + 2 * T * lds + 2 * T) : (3 * T * ld + T * lds + T));
}
extern "C" int qn_small_attn_fits(int T, int D) {
  return (D % 8 == 0 && D <= 128 && T >= 1 && sattn_smem(T, D, true) <= 200 * 1024) ? 1 : 0;
}
extern "C" int qn_small_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int T, int D,
                                 const long long* strides /* q,k,v,o: sb,sh,st each (12 values) */, float scale, int causal,
                                 cudaStream_t s) {
  using namespace qn::sattn;
  if (!qn_small_attn_fits(T, D)) return 970;
  Params p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.o = (__nv_bfloat16*)o;
  p.q_sb = strides[0]; p.q_sh = strides[1]; p.q_st = strides[2]; p.k_sb = strides[3]; p.k_sh = strides[4]; p.k_st = strides[5];
  p.v_sb = strides[6]; p.v_sh = strides[7]; p.v_st = strides[8]; p.o_sb = strides[9]; p.o_sh = strides[10]; p.o_st = strides[11];
  p.lse = lse; p.B = B; p.H = H; p.T = T; p.D = D; p.scale = scale; p.causal = causal;
  const size_t smem = sattn_smem(T, D, false);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(small_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  small_attn_fwd_kernel<<<dim3(H, B), THREADS, smem, s>>>(p);
  return (int)cudaGetLastError();
}
extern "C" int qn_small_attn_bwd(const void* q, const void* k, const void* v, const void* d_o, void* dq, void* dk, void* dv,
                                 int B, int H, int T, int D, const long long* strides /* q,k,v,do (12 values) */, float scale,
                                 int causal, cudaStream_t s) {
  using namespace qn::sattn;
  if (!qn_small_attn_fits(T, D)) return 970;
  Params p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.d_o = (const __nv_bfloat16*)d_o;
  p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
  p.q_sb = strides[0]; p.q_sh = strides[1]; p.q_st = strides[2]; p.k_sb = strides[3]; p.k_sh = strides[4]; p.k_st = strides[5];
  p.v_sb = strides[6]; p.v_sh = strides[7]; p.v_st = strides[8]; p.do_sb = strides[9]; p.do_sh = strides[10]; p.do_st = strides[11];
  p.B = B; p.H = H; p.T = T; p.D = D; p.scale = scale; p.causal = causal;
  const size_t smem = sattn_smem(T, D, true);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(small_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  small_attn_bwd_kernel<<<dim3(H, B), THREADS, smem, s>>>(p);
  return (int)cudaGetLastError();
}
