// quintnet_b200 — flash attention forward / backward on tcgen05 + TMEM + TMA (sm_100a), head_dim = 64 (K5).
//
// Works directly on the PACKED projection output  qkv[B, T, 3, H, 64]  (what the c_attn GEMM writes) and
// produces  out[B, T, H, 64]  (what the c_proj GEMM reads); backward writes the packed  dqkv[B, T, 3, H, 64].
// No head-split / transpose / concat copies exist anywhere around the kernel.
//
// forward  : one CTA per (128-query block, head, batch), 2 CTAs / SM.
//            warps 0-3 softmax (thread == query row == TMEM lane), warp 4 TMA producer, warp 5 MMA issuer.
//            S = Q·Kᵀ -> TMEM;  online softmax in registers (exp2, fp32);  P (bf16, dropout applied) -> smem;
//            P·V -> TMEM -> rescaled accumulation in registers;  O / l -> global,  LSE (log2 domain) saved.
// backward : one CTA per (128-key block, head, batch); loops over the query blocks that see it.
//            S = Q·Kᵀ and dP = dO·Vᵀ -> TMEM;  P, dS recomputed in registers -> smem (bf16);
//            dV += Pᵀ·dO, dK += dSᵀ·Q accumulate in TMEM across the loop;  dQ = dS·K -> TMEM -> fp32 atomics.
// All five backward GEMMs reuse the two smem tiles through K-major / MN-major UMMA descriptors.
#include "common.cuh"

#include <mutex>

namespace qn {
namespace attn {

constexpr uint32_t D = 64;          // head dim
constexpr uint32_t BQ = 128;        // query rows per block
constexpr uint32_t BKV = 128;       // key rows per block
constexpr uint32_t TILE = 128 * 128;          // bytes of a [128 x 64] bf16 tile (16 KB)
constexpr uint32_t kThreads = 192;

struct Params {
  uint32_t B, H, T;
  float scale_log2;                  // softmax scale * log2(e)
  float scale;                       // softmax scale
  int causal;
  float p_drop;                      // dropout probability (0 = off)
  uint32_t drop_thr16;               // keep iff hash16 >= thr
  float drop_scale;                  // 1 / (1 - p)
  uint64_t seed;
  __nv_bfloat16* out;                // fwd: [B,T,H,64]
  float* lse;                        // [B,H,T]  (log2 domain: m + log2(l))
  const float* delta;                // bwd: [B,H,T] rowsum(dO * O)
  float* dq_acc;                     // bwd: fp32 [B,T,H,64]
  __nv_bfloat16* dqkv;               // bwd: [B,T,3,H,64] (K and V slots written here)
};

QN_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                           int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// cheap counter hash -> two 16-bit uniforms (dropout mask for an element pair); identical in fwd and bwd
QN_DEVICE uint32_t hash2(uint32_t x, uint32_t key) {
  x ^= key;
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
QN_DEVICE uint32_t drop_key(uint64_t seed, uint32_t b, uint32_t h) {
  uint32_t k = static_cast<uint32_t>(seed) ^ (static_cast<uint32_t>(seed >> 32) * 0x9E3779B9u);
  k ^= (b * 0x85ebca6bu) ^ (h * 0xc2b2ae35u);
  k ^= k >> 13; k *= 0x5bd1e995u; k ^= k >> 15;
  return k;
}
// 2 x 16-bit uniforms for elements (q, k) and (q, k+1), k even (one multiply-xorshift round pair)
QN_DEVICE uint32_t drop_pair(uint32_t key, uint32_t idx) {
  uint32_t x = (idx ^ key) * 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  return x;
}
// AND-mask for a packed bf16 pair: keep element iff its 16-bit uniform >= thr
QN_DEVICE uint32_t drop_mask_pair(uint32_t h, uint32_t thr) {
  return (((h & 0xFFFFu) >= thr) ? 0x0000FFFFu : 0u) | (((h >> 16) >= thr) ? 0xFFFF0000u : 0u);
}

// write 8 packed bf16 (16 B) of row `r` at 16-byte chunk index `c16` (0..15 over 128 columns) into a
// [2 atoms][128 rows][128 B] K-major SW128 tile
QN_DEVICE void st_tile_chunk(uint8_t* tile, uint32_t r, uint32_t c16, uint4 v) {
  uint8_t* p = tile + (c16 >> 3) * TILE + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = v;
}

constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128, 0, 0);    // Q·Kᵀ, dO·Vᵀ : both K-major, N = 128
constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 64, 0, 1);    // P·V, dS·K    : A K-major, B MN-major, N = 64
constexpr uint32_t IDESC_TT = make_idesc_bf16(128, 64, 1, 1);    // Pᵀ·dO, dSᵀ·Q : both MN-major, N = 64

// ======================================================================================= forward
struct FwdSmem {
  static constexpr uint32_t Q = 0, K = TILE, V = 3 * TILE, P = 5 * TILE, BAR = 7 * TILE;
  static constexpr uint32_t BYTES = 7 * TILE + 256;   // 2 CTAs/SM: 2 x (BYTES + 1 KB reserved) <= 228 KB
};

// work item -> (query block, head, batch).  Items are ordered longest-first (large causal query blocks first)
// so the persistent CTAs finish together.
QN_DEVICE void fwd_item(uint32_t item, const Params& p, uint32_t nq, uint32_t& qb, uint32_t& h, uint32_t& b) {
  const uint32_t per_q = p.B * p.H;
  qb = nq - 1 - item / per_q;
  const uint32_t r = item % per_q;
  h = r % p.H;
  b = r / p.H;
}

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];      // SW128 tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;      // count 4 (one per softmax warp)
  uint64_t* o_full = bars + 7;
  uint64_t* q_empty = bars + 8;     // Q tile consumed by the last S MMA of a work item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  if ((smem_u32(smem) & 1023u) != 0 && threadIdx.x == 0) { printf("[quintnet_b200] attention smem misaligned\n"); __trap(); }
  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t nq = (p.T + BQ - 1) / BQ;
  const uint32_t nkv_total = (p.T + BKV - 1) / BKV;
  const uint32_t num_items = nq * p.H * p.B;

  if (warp == 4 && lane == 0) {
    prefetch_tensormap(&tmap);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer (persistent: prefetches across work items)
    // (the whole warp walks the loops with warp-uniform values; one elected lane issues — see the GEMM's MMA issuer)
    {
      uint32_t kvc = 0, nitem = 0;                       // running K/V stage counter, items done
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++nitem) {
        uint32_t qb, h, b;
        fwd_item(item, p, nq, qb, h, b);
        const uint32_t nkv = p.causal ? min(qb + 1, nkv_total) : nkv_total;
        mbar_wait(q_empty, (nitem & 1) ^ 1);             // previous item's S MMAs are done with the Q tile
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, TILE);
          tma_load_4d(smem + FwdSmem::Q, &tmap, q_full, 0, h, qb * BQ, b);                 // slot 0*H + h
        }
        __syncwarp();
        for (uint32_t j = 0; j < nkv; ++j, ++kvc) {
          const uint32_t s = kvc & 1, ph = (kvc >> 1) & 1;
          mbar_wait(&kv_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[s], 2 * TILE);
            tma_load_4d(smem + FwdSmem::K + s * TILE, &tmap, &kv_full[s], 0, p.H + h, j * BKV, b);
            tma_load_4d(smem + FwdSmem::V + s * TILE, &tmap, &kv_full[s], 0, 2 * p.H + h, j * BKV, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------- MMA issuer
    // Whole warp, warp-uniform control flow; only the tcgen05 instructions sit behind elect_one().  (Issuing from a
    // divergent `lane == 0` region made the compiler wrap every MMA in an ELECT / R2UR / branch waterfall of ~40 SASS
    // instructions: 12 MMAs per key block were paced by a single thread's instruction latency, not by the tensor pipe.)
    {
      const uint32_t sQ = smem_u32(smem + FwdSmem::Q), sP = smem_u32(smem + FwdSmem::P);
      const uint64_t dK0 = make_smem_desc_sw128(0, 1024, 0), dMN0 = make_smem_desc_sw128(0, 1024, TILE);
      auto kdesc = [&](uint32_t addr) { return dK0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
      auto mndesc = [&](uint32_t addr) { return dMN0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
      uint32_t kvc = 0, itc = 0, nitem = 0;
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++nitem) {
        uint32_t qb, h, b;
        fwd_item(item, p, nq, qb, h, b);
        const uint32_t nkv = p.causal ? min(qb + 1, nkv_total) : nkv_total;
        mbar_wait(q_full, nitem & 1);
        for (uint32_t j = 0; j < nkv; ++j, ++kvc, ++itc) {
          const uint32_t s = kvc & 1, ph = (kvc >> 1) & 1;
          const uint32_t sK = smem_u32(smem + FwdSmem::K + s * TILE), sV = smem_u32(smem + FwdSmem::V + s * TILE);
          mbar_wait(&kv_full[s], ph);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < D / 16; ++k)
              umma_bf16_ss(tmem_S, kdesc(sQ + k * 32), kdesc(sK + k * 32), IDESC_S, k != 0);
            umma_commit(s_full);
            if (j + 1 == nkv) umma_commit(q_empty);      // Q may be overwritten by the next item's load
          }
          __syncwarp();
          mbar_wait(p_full, itc & 1);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < BKV / 16; ++k)
              umma_bf16_ss(tmem_O, kdesc(sP + (k >> 2) * TILE + (k & 3) * 32), mndesc(sV + k * 2048), IDESC_PV, k != 0);
            umma_commit(&kv_empty[s]);
            umma_commit(o_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulate warps
    const uint32_t r = warp * 32 + lane;               // row within the query block == TMEM lane
    const uint32_t lane_base = (warp * 32u) << 16;
    uint8_t* sP = smem + FwdSmem::P;
    uint32_t itc = 0;
    for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x) {
      uint32_t qb, h, b;
      fwd_item(item, p, nq, qb, h, b);
      const uint32_t nkv = p.causal ? min(qb + 1, nkv_total) : nkv_total;
      const uint32_t q = qb * BQ + r;
      const uint32_t dkey = drop_key(p.seed, b, h);
      float m = -INFINITY, l = 0.f;
      float o[D];
#pragma unroll
      for (int e = 0; e < (int)D; ++e) o[e] = 0.f;
      for (uint32_t j = 0; j < nkv; ++j, ++itc) {
        const uint32_t k0 = j * BKV;
        const bool need_mask = (p.causal && j == qb) || (k0 + BKV > p.T);
        mbar_wait(s_full, itc & 1);
        tcgen05_fence_after();
        // pass 1: row max of the RAW scores (scale > 0 commutes with max); masked variant only on the
        // diagonal / tail block (warp-uniform branch outside the element loops)
        float mx_raw = -INFINITY;
        if (!need_mask) {
#pragma unroll 1
          for (uint32_t c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) mx_raw = fmaxf(mx_raw, __uint_as_float(v[e]));
          }
        } else {
          const uint32_t klim = min(p.T, p.causal ? q + 1 : p.T);    // valid keys: kk < klim
#pragma unroll 1
          for (uint32_t c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (k0 + c * 32 + e < klim) mx_raw = fmaxf(mx_raw, __uint_as_float(v[e]));
          }
        }
        const float m_new = fmaxf(m, mx_raw * p.scale_log2);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;      // fully-masked row guard
        const float alpha = exp2f(m - m_use);                        // m = -inf -> 0
        float lsum = 0.f;
        // pass 2: P = exp2(s*scale - m), rowsum, (dropout as an AND-mask on the packed bf16), -> smem
        const uint32_t klim2 = need_mask ? min(p.T, p.causal ? q + 1 : p.T) : 0xFFFFFFFFu;
        const bool drop = p.p_drop > 0.f;
        const uint32_t ridx = q * p.T + k0;
#pragma unroll 1
        for (uint32_t c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (!need_mask) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float p0 = exp2f(fmaf(__uint_as_float(v[e]), p.scale_log2, -m_use));
              const float p1 = exp2f(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, -m_use));
              lsum += p0 + p1;
              pk[e >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const uint32_t kk = k0 + c * 32 + e;
              const float p0 = (kk < klim2) ? exp2f(fmaf(__uint_as_float(v[e]), p.scale_log2, -m_use)) : 0.f;
              const float p1 = (kk + 1 < klim2) ? exp2f(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, -m_use)) : 0.f;
              lsum += p0 + p1;
              pk[e >> 1] = pack_bf16(p0, p1);
            }
          }
          if (drop) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] &= drop_mask_pair(drop_pair(dkey, ridx + c * 32 + 2 * e), p.drop_thr16);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            st_tile_chunk(sP, r, c * 4 + g, make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]));
        }
        l = l * alpha + lsum;
        m = m_new;
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // P·V result -> rescaled accumulation
        mbar_wait(o_full, itc & 1);
        tcgen05_fence_after();
#pragma unroll
        for (uint32_t c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_O + lane_base + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * alpha + __uint_as_float(v[e]);
        }
        tcgen05_fence_before();
      }
      if (q < p.T) {
        const float inv = l > 0.f ? p.drop_scale / l : 0.f;     // dropout 1/(1-p) folded here
        __nv_bfloat16* op = p.out + ((static_cast<size_t>(b) * p.T + q) * p.H + h) * D;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 pk;
          pk.x = pack_bf16(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv); pk.y = pack_bf16(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
          pk.z = pack_bf16(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv); pk.w = pack_bf16(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
          *reinterpret_cast<uint4*>(op + g * 8) = pk;
        }
        p.lse[(static_cast<size_t>(b) * p.H + h) * p.T + q] = (l > 0.f) ? m + log2f(l) : -INFINITY;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ======================================================================================= backward
// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                  uint32_t B, uint32_t H, uint32_t T) {
  const size_t row = static_cast<size_t>(blockIdx.x) * 32 + (threadIdx.x >> 3);   // (b, t, h) flattened
  const uint32_t part = threadIdx.x & 7;                                           // 8 threads x 8 elements
  const size_t total = static_cast<size_t>(B) * T * H;
  float s = 0.f;
  if (row < total) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * D + part * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + row * D + part * 8);
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
    s = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (part == 0 && row < total) {
    const uint32_t hh = row % H;
    const size_t bt = row / H;
    const uint32_t t = bt % T, b = bt / T;
    delta[(static_cast<size_t>(b) * H + hh) * T + t] = s;
  }
}

struct BwdSmem {
  // K/V double-buffered across work items, Q/dO 2-stage ring, P and dS tiles
  static constexpr uint32_t K = 0, V = 2 * TILE, Q = 4 * TILE, DO = 6 * TILE, P = 8 * TILE, DS = 10 * TILE, BAR = 12 * TILE;
  static constexpr uint32_t BYTES = 12 * TILE + 256;
};

constexpr uint32_t kBwdThreads = 320;   // 8 compute warps + TMA warp + MMA warp

// work item -> (key block, head, batch); causal work per item shrinks with kb, so ascending kb = longest first
QN_DEVICE void bwd_item(uint32_t item, const Params& p, uint32_t& kb, uint32_t& h, uint32_t& b) {
  const uint32_t per_k = p.B * p.H;
  kb = item / per_k;
  const uint32_t r = item % per_k;
  h = r % p.H;
  b = r / p.H;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BAR);
  uint64_t* kv_full = bars;         // [2]
  uint64_t* kv_empty = bars + 2;    // [2]
  uint64_t* q_full = bars + 4;      // [2]
  uint64_t* q_empty = bars + 6;     // [2]
  uint64_t* sdp_full = bars + 8;
  uint64_t* pds_full = bars + 9;    // count 8 (one per compute warp)
  uint64_t* dq_full = bars + 10;    // [2]  (dQ accumulator is double-buffered in TMEM)
  uint64_t* done = bars + 12;       // all MMAs of a work item retired (dK / dV final)
  uint64_t* dkdv_empty = bars + 13; // count 8: dK / dV accumulators read out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  if ((smem_u32(smem) & 1023u) != 0 && threadIdx.x == 0) { printf("[quintnet_b200] attention smem misaligned\n"); __trap(); }
  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t nq_total = (p.T + BQ - 1) / BQ;
  const uint32_t nkb = (p.T + BKV - 1) / BKV;
  const uint32_t num_items = nkb * p.H * p.B;
  constexpr uint32_t kTmaWarp = 8, kMmaWarp = 9;

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tensormap(&tmap_qkv);
    prefetch_tensormap(&tmap_do);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&dq_full[i], 1);
    }
    mbar_init(sdp_full, 1); mbar_init(pds_full, 8); mbar_init(done, 1); mbar_init(dkdv_empty, 8);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // TMEM columns: S 0-127 | dP 128-255 | dV 256-319 | dK 320-383 | dQ[0] 384-447 | dQ[1] 448-511
  const uint32_t tmem_S = tmem, tmem_dP = tmem + 128, tmem_dV = tmem + 256, tmem_dK = tmem + 320, tmem_dQ = tmem + 384;

  if (warp == kTmaWarp) {
    {   // whole warp, one elected lane issues
      uint32_t qc = 0, nitem = 0;
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++nitem) {
        uint32_t kb, h, b;
        bwd_item(item, p, kb, h, b);
        const uint32_t qb_first = p.causal ? kb : 0, niter = nq_total - qb_first;
        const uint32_t kvb = nitem & 1, kph = (nitem >> 1) & 1;
        mbar_wait(&kv_empty[kvb], kph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[kvb], 2 * TILE);
          tma_load_4d(smem + BwdSmem::K + kvb * TILE, &tmap_qkv, &kv_full[kvb], 0, p.H + h, kb * BKV, b);
          tma_load_4d(smem + BwdSmem::V + kvb * TILE, &tmap_qkv, &kv_full[kvb], 0, 2 * p.H + h, kb * BKV, b);
        }
        __syncwarp();
        for (uint32_t it = 0; it < niter; ++it, ++qc) {
          const uint32_t s = qc & 1, ph = (qc >> 1) & 1;
          const uint32_t q0 = (qb_first + it) * BQ;
          mbar_wait(&q_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[s], 2 * TILE);
            tma_load_4d(smem + BwdSmem::Q + s * TILE, &tmap_qkv, &q_full[s], 0, h, q0, b);
            tma_load_4d(smem + BwdSmem::DO + s * TILE, &tmap_do, &q_full[s], 0, h, q0, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kMmaWarp) {
    {   // whole warp, warp-uniform control flow; tcgen05 instructions behind elect_one() (see the forward kernel)
      const uint32_t sP = smem_u32(smem + BwdSmem::P), sDS = smem_u32(smem + BwdSmem::DS);
      const uint64_t dK0 = make_smem_desc_sw128(0, 1024, 0), dMN0 = make_smem_desc_sw128(0, 1024, TILE);
      auto kdesc = [&](uint32_t addr) { return dK0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
      auto mndesc = [&](uint32_t addr) { return dMN0 | static_cast<uint64_t>((addr & 0x3FFFF) >> 4); };
      uint32_t qc = 0, itc = 0, nitem = 0;
      for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++nitem) {
        uint32_t kb, h, b;
        bwd_item(item, p, kb, h, b);
        const uint32_t qb_first = p.causal ? kb : 0, niter = nq_total - qb_first;
        const uint32_t kvb = nitem & 1;
        const uint32_t sK = smem_u32(smem + BwdSmem::K + kvb * TILE), sV = smem_u32(smem + BwdSmem::V + kvb * TILE);
        auto wait_q = [&](uint32_t q_count) {         // all lanes: the query block in ring slot q_count has landed
          mbar_wait(&q_full[q_count & 1], (q_count >> 1) & 1);
          tcgen05_fence_after();
        };
        auto issue_s_dp = [&](uint32_t q_count) {     // elected lane: S = Q·Kᵀ, dP = dO·Vᵀ of that query block
          const uint32_t s = q_count & 1;
          const uint32_t sQ = smem_u32(smem + BwdSmem::Q + s * TILE), sDO = smem_u32(smem + BwdSmem::DO + s * TILE);
#pragma unroll
          for (uint32_t k = 0; k < D / 16; ++k)
            umma_bf16_ss(tmem_S, kdesc(sQ + k * 32), kdesc(sK + k * 32), IDESC_S, k != 0);
#pragma unroll
          for (uint32_t k = 0; k < D / 16; ++k)
            umma_bf16_ss(tmem_dP, kdesc(sDO + k * 32), kdesc(sV + k * 32), IDESC_S, k != 0);
          umma_commit(sdp_full);
        };
        mbar_wait(&kv_full[kvb], (nitem >> 1) & 1);
        wait_q(qc);
        if (elect_one()) issue_s_dp(qc);
        __syncwarp();
        for (uint32_t it = 0; it < niter; ++it, ++qc, ++itc) {
          const uint32_t s = qc & 1;
          const uint32_t sQ = smem_u32(smem + BwdSmem::Q + s * TILE), sDO = smem_u32(smem + BwdSmem::DO + s * TILE);
          mbar_wait(pds_full, itc & 1);                // P / dS of this block are in smem, S / dP TMEM consumed
          tcgen05_fence_after();
          // the next block's S / dP go FIRST so the compute warps can start on them while the tensor pipe
          // is busy with this block's dV / dK / dQ
          if (it + 1 < niter) wait_q(qc + 1);
          if (it == 0 && nitem > 0) {                  // previous item's dK / dV must have been read out
            mbar_wait(dkdv_empty, (nitem - 1) & 1);
            tcgen05_fence_after();
          }
          if (elect_one()) {
            if (it + 1 < niter) issue_s_dp(qc + 1);
#pragma unroll
            for (uint32_t k = 0; k < BQ / 16; ++k)     // dV += Pᵀ · dO   (A = P as MN-major, B = dO as MN-major)
              umma_bf16_ss(tmem_dV, mndesc(sP + k * 2048), mndesc(sDO + k * 2048), IDESC_TT, (it | k) != 0);
#pragma unroll
            for (uint32_t k = 0; k < BQ / 16; ++k)     // dK += dSᵀ · Q
              umma_bf16_ss(tmem_dK, mndesc(sDS + k * 2048), mndesc(sQ + k * 2048), IDESC_TT, (it | k) != 0);
#pragma unroll
            for (uint32_t k = 0; k < BKV / 16; ++k)    // dQ = dS · K     (A = dS K-major, B = K as MN-major)
              umma_bf16_ss(tmem_dQ + (itc & 1) * 64, kdesc(sDS + (k >> 2) * TILE + (k & 3) * 32), mndesc(sK + k * 2048),
                           IDESC_PV, k != 0);
            umma_commit(&q_empty[s]);
            umma_commit(&dq_full[itc & 1]);
            if (it + 1 == niter) {
              umma_commit(&kv_empty[kvb]);
              umma_commit(done);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ 8 compute warps
    const uint32_t qtr = warp & 3, half = warp >> 2;          // TMEM lane quarter, column half
    const uint32_t r = qtr * 32 + lane;
    const uint32_t lane_base = (qtr * 32u) << 16;
    uint8_t* sP = smem + BwdSmem::P;
    uint8_t* sDS = smem + BwdSmem::DS;
    const float c_ds = p.scale * p.drop_scale;
    const bool drop = p.p_drop > 0.f;
    uint32_t itc = 0, nitem = 0;
    for (uint32_t item = blockIdx.x; item < num_items; item += gridDim.x, ++nitem) {
      uint32_t kb, h, b;
      bwd_item(item, p, kb, h, b);
      const uint32_t k0 = kb * BKV;
      const uint32_t qb_first = p.causal ? kb : 0, niter = nq_total - qb_first;
      const uint32_t dkey = drop_key(p.seed, b, h);
      auto drain_dq = [&](uint32_t c_it, uint32_t qblk) {      // dQ block -> fp32 atomics (32 columns per thread)
        const uint32_t q = qblk * BQ + r;
        mbar_wait(&dq_full[c_it & 1], (c_it >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_dQ + (c_it & 1) * 64 + lane_base + half * 32, v);
        tmem_ld_wait();
        if (q < p.T) {
          float* dq = p.dq_acc + ((static_cast<size_t>(b) * p.T + q) * p.H + h) * D + half * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dq + e), "f"(__uint_as_float(v[e])),
                         "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                         : "memory");
        }
        tcgen05_fence_before();
      };
      // row statistics are software-pipelined one query block ahead: the loads of block it+1 are issued
      // before the barrier waits of block it+1, so their DRAM latency never sits on the critical path
      const float* lse_bh = p.lse + (static_cast<size_t>(b) * p.H + h) * p.T;
      const float* dl_bh = p.delta + (static_cast<size_t>(b) * p.H + h) * p.T;
      const float inv_rs = 1.0f / p.drop_scale;
      float lse_nx, dl_nx;
      {
        const uint32_t q1 = qb_first * BQ + r;
        lse_nx = q1 < p.T ? lse_bh[q1] : INFINITY;               // rows beyond T -> p = 0
        dl_nx = q1 < p.T ? dl_bh[q1] * inv_rs : 0.f;
      }
      for (uint32_t it = 0; it < niter; ++it, ++itc) {
        const uint32_t qb = qb_first + it;
        const uint32_t q = qb * BQ + r;
        const float lse = lse_nx, dl2 = dl_nx;
        if (it + 1 < niter) {
          const uint32_t q1 = q + BQ;
          lse_nx = q1 < p.T ? lse_bh[q1] : INFINITY;
          dl_nx = q1 < p.T ? dl_bh[q1] * inv_rs : 0.f;
        }
        const bool need_mask = (p.causal && qb == kb) || (k0 + BKV > p.T);
        const uint32_t klim = need_mask ? min(p.T, p.causal ? q + 1 : p.T) : 0xFFFFFFFFu;
        const uint32_t ridx = q * p.T + k0;
        mbar_wait(sdp_full, itc & 1);
        tcgen05_fence_after();
        // dS = (P·scale·rs) ⊙ (mask·dP − δ/rs);  dV uses mask·P (rs applied when dV is written)
#pragma unroll 1
        for (uint32_t cc = 0; cc < 2; ++cc) {
          const uint32_t c = half * 2 + cc;                      // this warp's 32-column chunk
          uint32_t sv[32], dv[32];
          tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, sv);
          tmem_ld_32x32b_x32(tmem_dP + lane_base + c * 32, dv);
          tmem_ld_wait();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = exp2f(fmaf(__uint_as_float(sv[e]), p.scale_log2, -lse));
            float p1 = exp2f(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -lse));
            if (need_mask) {
              const uint32_t kk = k0 + c * 32 + e;
              p0 = (kk < klim) ? p0 : 0.f;
              p1 = (kk + 1 < klim) ? p1 : 0.f;
            }
            float d0 = __uint_as_float(dv[e]), d1 = __uint_as_float(dv[e + 1]);
            uint32_t pw = pack_bf16(p0, p1);
            if (drop) {
              const uint32_t h2 = drop_pair(dkey, ridx + c * 32 + e);
              const bool k0b = (h2 & 0xFFFFu) >= p.drop_thr16, k1b = (h2 >> 16) >= p.drop_thr16;
              d0 = k0b ? d0 : 0.f;
              d1 = k1b ? d1 : 0.f;
              pw &= (k0b ? 0x0000FFFFu : 0u) | (k1b ? 0xFFFF0000u : 0u);
            }
            pk[e >> 1] = pw;
            dk[e >> 1] = pack_bf16((p0 * c_ds) * (d0 - dl2), (p1 * c_ds) * (d1 - dl2));
          }
          // the P / dS smem tiles are still being read by the dV / dK / dQ MMAs of the previous query block:
          // the register math above overlapped with them, the stores must wait for their completion
          if (cc == 0 && itc > 0) mbar_wait(&dq_full[(itc - 1) & 1], ((itc - 1) >> 1) & 1);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            st_tile_chunk(sP, r, c * 4 + g, make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]));
            st_tile_chunk(sDS, r, c * 4 + g, make_uint4(dk[g * 4 + 0], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]));
          }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        if (it > 0) drain_dq(itc - 1, qb - 1);       // dQ of the previous block: TMEM -> fp32 atomics
      }
      drain_dq(itc - 1, qb_first + niter - 1);
      // dK / dV of this key block (thread == key row): half 0 writes dK, half 1 writes dV
      mbar_wait(done, nitem & 1);
      tcgen05_fence_after();
      const uint32_t kk = k0 + r;
      {
        const uint32_t which = half;                           // 0 -> dK (slot 1), 1 -> dV (slot 2)
        const uint32_t taddr = (which == 0 ? tmem_dK : tmem_dV) + lane_base;
        const float osc = (which == 1) ? p.drop_scale : 1.0f;  // dV carries the dropout 1/(1-p)
        __nv_bfloat16* dst = p.dqkv + (((static_cast<size_t>(b) * p.T + (kk < p.T ? kk : 0)) * 3 + (which + 1)) * p.H + h) * D;
#pragma unroll
        for (uint32_t c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (kk < p.T) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 pk;
              pk.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * osc, __uint_as_float(v[g * 8 + 1]) * osc);
              pk.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * osc, __uint_as_float(v[g * 8 + 3]) * osc);
              pk.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * osc, __uint_as_float(v[g * 8 + 5]) * osc);
              pk.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * osc, __uint_as_float(v[g * 8 + 7]) * osc);
              *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkdv_empty);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// dq_acc fp32 [B,T,H,64] -> dqkv slot 0 (bf16)
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, uint32_t B, uint32_t H,
                       uint32_t T) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;      // one 8-element group
  const size_t total = static_cast<size_t>(B) * T * H * (D / 8);
  if (i >= total) return;
  const uint32_t g = i % (D / 8);
  const size_t row = i / (D / 8);                 // (b, t, h)
  const uint32_t hh = row % H;
  const size_t bt = row / H;
  const float4 a = *reinterpret_cast<const float4*>(dq_acc + row * D + g * 8);
  const float4 c = *reinterpret_cast<const float4*>(dq_acc + row * D + g * 8 + 4);
  uint4 pk;
  pk.x = pack_bf16(a.x, a.y); pk.y = pack_bf16(a.z, a.w); pk.z = pack_bf16(c.x, c.y); pk.w = pack_bf16(c.z, c.w);
  *reinterpret_cast<uint4*>(dqkv + ((bt * 3 + 0) * H + hh) * D + g * 8) = pk;
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// tensor [B, T, S, 64] bf16 (S = 3H for the packed qkv, H for dO); box = {64, 1, 128, 1}
static int make_tmap4(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t T, uint64_t S) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 901;
  cuuint64_t dims[4] = {D, S, T, B};
  cuuint64_t strides[3] = {D * 2, S * D * 2, T * S * D * 2};
  cuuint32_t box[4] = {D, 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // first CUDA call on this host thread (e.g. an autograd worker): bind the primary context and retry
    cudaFree(nullptr);
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? 0 : 900 + static_cast<int>(r);
}

static void fill_params(Params& p, int B, int H, int T, float scale, int causal, float p_drop, unsigned long long seed) {
  p.B = B; p.H = H; p.T = T;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.p_drop = p_drop;
  p.drop_thr16 = static_cast<uint32_t>(p_drop * 65536.0f);
  p.drop_scale = p_drop > 0.f ? 1.0f / (1.0f - static_cast<float>(p.drop_thr16) / 65536.0f) : 1.0f;
  p.seed = seed;
}

}  // namespace attn
}  // namespace qn

// qkv: bf16 [B,T,3,H,64] contiguous; out: bf16 [B,T,H,64]; lse: fp32 [B,H,T]
extern "C" int qn_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int T, float scale, int causal,
                           float p_drop, unsigned long long seed, cudaStream_t s) {
  using namespace qn::attn;
  CUtensorMap tm;
  if (int rc = make_tmap4(&tm, qkv, B, T, 3 * H)) return rc;
  Params p{};
  fill_params(p, B, H, T, scale, causal, p_drop, seed);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned items = static_cast<unsigned>((T + BQ - 1) / BQ) * H * B;
  const unsigned grid = items < 2u * num_sms ? items : 2u * num_sms;      // persistent: 2 CTAs per SM
  attn_fwd_kernel<<<grid, kThreads, FwdSmem::BYTES, s>>>(tm, p);
  return (int)cudaGetLastError();
}

// dout/out: bf16 [B,T,H,64]; dqkv: bf16 [B,T,3,H,64] (fully written); workspace: fp32 [B*T*H*64 + B*H*T]
extern "C" int qn_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           float* workspace, int B, int H, int T, float scale, int causal, float p_drop,
                           unsigned long long seed, cudaStream_t s) {
  using namespace qn::attn;
  CUtensorMap tq, tdo;
  if (int rc = make_tmap4(&tq, qkv, B, T, 3 * H)) return rc;
  if (int rc = make_tmap4(&tdo, dout, B, T, H)) return rc;
  Params p{};
  fill_params(p, B, H, T, scale, causal, p_drop, seed);
  const size_t n_dq = static_cast<size_t>(B) * T * H * D;
  float* dq_acc = workspace;
  float* delta = workspace + n_dq;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  cudaMemsetAsync(dq_acc, 0, n_dq * sizeof(float), s);
  const size_t rows = static_cast<size_t>(B) * T * H;
  attn_delta_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(out),
                                                                 static_cast<const __nv_bfloat16*>(dout), delta, B, H, T);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned items = static_cast<unsigned>((T + BKV - 1) / BKV) * H * B;
  const unsigned grid = items < static_cast<unsigned>(num_sms) ? items : num_sms;   // persistent: 1 CTA per SM
  attn_bwd_kernel<<<grid, kBwdThreads, BwdSmem::BYTES, s>>>(tq, tdo, p);
  const size_t groups = rows * (D / 8);
  attn_dq_convert_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, s>>>(dq_acc, p.dqkv, B, H, T);
  return (int)cudaGetLastError();
}
