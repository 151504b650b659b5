// quintnet_b200 — block-scaled FP8 GEMM (OCP MXFP8: e4m3 elements, one UE8M0 power-of-two scale per 32 elements
// along K) on tcgen05.mma.kind::mxf8f6f4.block_scale (sm_100a).
//
//   C[M, N] (bf16) = ( A8 ⊙ 2^SFA ) · ( B8 ⊙ 2^SFB )ᵀ (+ bias),   fp32 accumulation in tensor memory.
//
// The scale factors never touch the CUDA cores: they travel global -> shared memory with the operand tiles (bulk async
// copies on the same mbarrier), shared -> tensor memory with tcgen05.cp, and the tensor core applies them per 32-element
// K block inside the MMA.  (The round-1 FP8 path, qn_gemm_fp8 in gemm.cu, scales per ROW and multiplies the scales back
// in the epilogue; this is the block-scaled format the Blackwell tensor core implements natively.)
//
// Layout contract (produced by qn_quantize_mxfp8 below, mirrored by ops/cuda_ops.py::quantize_mx_reference):
//   data   : e4m3 bytes, row-major [rows, K], K % 128 == 0, leading dimension in bytes (multiple of 16)
//   scales : UE8M0 bytes (value 2^(b - 127)) in 512-byte ATOMS, one per (128-row block, 128-wide K block):
//            atom index = row_blk * (K / 128) + k_blk ; byte offset inside = (r % 32) * 16 + (r / 32) * 4 + j  for row r of
//            the block and 32-element K block j (0..3) — exactly the image tcgen05.cp.32x128b.warpx4 moves into 4 TMEM
//            columns (lane r % 32 of every sub-partition, column r / 32, byte j), which is where the MMA reads them.
//
// Kernel structure (one CTA per SM, 256 threads, persistent over 128 x 128 output tiles):
//   warp 0 : TMA producer — A / B tiles [128 x 128 B] (128-byte swizzle) + the two 512-byte scale atoms, 5-stage ring
//   warp 1 : issuer — per stage: 2 x tcgen05.cp (scales -> TMEM), 4 x tcgen05.mma (K = 32 each, scale id 0..3), commit
//   warp 2 : TMEM allocator (512 columns: 2 x 128 accumulator columns + 8 scale columns per stage)
//   warps 4-7: epilogue — tcgen05.ld -> (+bias) -> bf16 -> swizzled smem -> TMA store
#include "common.cuh"

#include <cuda_fp8.h>

namespace qn {
namespace mx {

constexpr uint32_t BM = 128, BN = 128, BKB = 128;        // tile rows / cols / K bytes (= fp8 elements) per stage
constexpr uint32_t kStages = 5;
constexpr uint32_t A_BYTES = BM * BKB, B_BYTES = BN * BKB, SF_BYTES = 512;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + 2 * SF_BYTES;          // 33 KB
constexpr uint32_t EPI_BYTES = 4 * 2 * 4096;                               // 4 warps x two [32 x 64] bf16 slabs
constexpr uint32_t BAR_BYTES = 256;
constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
constexpr uint32_t kThreads = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SF_COL0 = 256;                                          // scale columns start after the accumulators

struct Params {
  uint32_t M, N, K;                  // K in elements (= bytes)
  const uint8_t* sfa; const uint8_t* sfb;
  const __nv_bfloat16* bias;
};

// E4M3 x E4M3, fp32 accumulate, UE8M0 scales every 32 elements (layout: cute::UMMA::InstrDescriptorBlockScaled)
//   [4,6) b_sf_id  [7,10) a_format=0  [10,13) b_format=0  [15] a_major=K  [16] b_major=K  [17,23) N>>3
//   [23] scale_format 1=E8M0  [24,29) M>>4  [29,31) a_sf_id  [31] k_size 0 (K = 32)
__host__ __device__ constexpr uint32_t make_idesc_mx(uint32_t M, uint32_t N, uint32_t sf_id) {
  return (sf_id << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (sf_id << 29);
}
// shared-memory descriptor of a 512-byte scale atom for tcgen05.cp: K-major, no swizzle, 8-row core matrices 128 B apart
QN_DEVICE uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 32;              // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (sm_100)
  return d;                                                // layout type 0 = no swizzle
}
QN_DEVICE void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
QN_DEVICE void umma_mxf8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                            uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + kStages * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN, num_tiles = tiles_m * tiles_n;
  const uint32_t num_kb = p.K / BKB;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a); prefetch_tensormap(&tmap_b); prefetch_tensormap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ producer (whole warp walks, one lane issues)
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const uint32_t m_blk = tile % tiles_m, n_blk = tile / tiles_m;     // m fastest: CTAs of a wave share the B tile
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BKB, m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tmap_b, &full_bar[stage], kb * BKB, n_blk * BN);
          bulk_load_1d(sa + A_BYTES + B_BYTES, p.sfa + (static_cast<size_t>(m_blk) * num_kb + kb) * SF_BYTES, SF_BYTES,
                       &full_bar[stage]);
          bulk_load_1d(sa + A_BYTES + B_BYTES + SF_BYTES, p.sfb + (static_cast<size_t>(n_blk) * num_kb + kb) * SF_BYTES,
                       SF_BYTES, &full_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ scale copy + MMA issuer
    const uint64_t d_hi = make_smem_desc_sw128(0, 1024, 0);
    const uint32_t smem_base = smem_u32(smem);
    uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint64_t adesc = d_hi | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
        const uint64_t bdesc = d_hi | static_cast<uint64_t>(((sa + A_BYTES) & 0x3FFFF) >> 4);
        const uint32_t t_sfa = tmem_base + SF_COL0 + stage * 8, t_sfb = t_sfa + 4;
        if (elect_one()) {
          tmem_cp_32x128b_warpx4(t_sfa, make_sf_desc(sa + A_BYTES + B_BYTES));
          tmem_cp_32x128b_warpx4(t_sfb, make_sf_desc(sa + A_BYTES + B_BYTES + SF_BYTES));
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k)                    // 4 x (K = 32 elements = 32 bytes = 2 descriptor units)
            umma_mxf8_ss(tmem_d, adesc + k * 2, bdesc + k * 2, make_idesc_mx(BM, BN, k), (kb | k) != 0 ? 1u : 0u, t_sfa, t_sfb);
          umma_commit(&empty_bar[stage]);
          if (kb + 1 == num_kb) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (4 warps = 128 accumulator lanes)
    const uint32_t q = warp - 4;
    uint8_t* sbuf0 = smem_epi + q * 8192;
    uint32_t as = 0, aphase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const uint32_t m_blk = tile % tiles_m, n_blk = tile / tiles_m;
      const uint32_t m0 = m_blk * BM + q * 32, n0 = n_blk * BN;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
#pragma unroll
      for (uint32_t chunk = 0; chunk < 2; ++chunk) {             // two [32 x 64] slabs per warp
        const uint32_t nc = n0 + chunk * 64;
        uint8_t* sbuf = sbuf0 + chunk * 4096;
        uint8_t* srow = sbuf + lane * 128;
        if (lane == 0) tma_store_wait_read<1>();                   // the store that last read this slab buffer is done
        __syncwarp();
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BN + chunk * 64;
#pragma unroll
        for (uint32_t hcol = 0; hcol < 2; ++hcol) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + hcol * 32, acc);
          tmem_ld_wait();
          float* f = reinterpret_cast<float*>(acc);
          const uint32_t nh = nc + hcol * 32;
          if (p.bias != nullptr) {
#pragma unroll
            for (uint32_t v = 0; v < 4; ++v) {
              if (nh + v * 8 < p.N) {
                const uint4 b = ld_nc_v4(p.bias + nh + v * 8);
                const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
                f[v * 8 + 0] += b0.x; f[v * 8 + 1] += b0.y; f[v * 8 + 2] += b1.x; f[v * 8 + 3] += b1.y;
                f[v * 8 + 4] += b2.x; f[v * 8 + 5] += b2.y; f[v * 8 + 6] += b3.x; f[v * 8 + 7] += b3.y;
              }
            }
          }
#pragma unroll
          for (uint32_t v = 0; v < 4; ++v) {
            uint4 pk;
            pk.x = pack_bf16(f[v * 8 + 0], f[v * 8 + 1]); pk.y = pack_bf16(f[v * 8 + 2], f[v * 8 + 3]);
            pk.z = pack_bf16(f[v * 8 + 4], f[v * 8 + 5]); pk.w = pack_bf16(f[v * 8 + 6], f[v * 8 + 7]);
            *reinterpret_cast<uint4*>(srow + (((hcol * 4 + v) ^ (lane & 7)) << 4)) = pk;      // 128B swizzle
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && nc < p.N) {
          tma_store_2d(&tmap_c, sbuf, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
          tma_store_commit();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ quantiser
// x [rows, K] bf16 (row stride ldx elements)  ->  e4m3 bytes [rows, K] (row stride ldq bytes) + UE8M0 scale atoms.
// OCP MX rule: shared exponent = floor(log2(amax of the 32-element block)) - 8 (8 = emax of e4m3), elements are
// x * 2^-e converted round-to-nearest-even with saturation to +-448.  One thread quantises 8 elements; 4 neighbouring
// lanes form a 32-element block.  Rows beyond `rows` (up to the next multiple of 128) get scale byte 0 / zero data
// only in the scale atoms (the data rows are never read: TMA zero-fills out-of-bounds rows).
__global__ void __launch_bounds__(256)
quantize_mxfp8_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf, int rows, int K,
                      long long ldx, long long ldq) {
  const int vec_per_row = K / 8;
  const long long nvec = static_cast<long long>(rows) * vec_per_row;
  const int num_kb = K / 128;
  const long long nvec_pad = (nvec + 31) / 32 * 32;                 // whole warps iterate: the shuffles below are warp-wide
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < nvec_pad;
       i0 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool valid = i0 < nvec;
    const long long i = valid ? i0 : nvec - 1;
    const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i % vec_per_row) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(x + r * ldx + c);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
    float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
    float amax = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(v[e]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));     // (vec_per_row % 4 == 0: the 4 lanes share a row)
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    int e8 = 0;                                                     // biased UE8M0 exponent
    if (amax > 0.f) {
      const int ex = static_cast<int>((__float_as_uint(amax) >> 23) & 0xFF) - 127;   // floor(log2(amax)) (normal numbers)
      e8 = min(max(ex - 8 + 127, 0), 254);
    }
    const float inv = __uint_as_float(static_cast<uint32_t>(254 - e8) << 23);         // 2^-(e8 - 127)
    uint2 o;
    uint16_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      h[e] = __nv_cvt_float2_to_fp8x2(make_float2(v[2 * e] * inv, v[2 * e + 1] * inv), __NV_SATFINITE, __NV_E4M3);
    o.x = h[0] | (static_cast<uint32_t>(h[1]) << 16);
    o.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
    if (!valid) continue;
    *reinterpret_cast<uint2*>(q + r * ldq + c) = o;
    if ((threadIdx.x & 3) == 0) {                                   // one lane of the 4 writes the block's scale
      const int kb = c / 128, j = (c % 128) / 32, rb = r / 128, rr = r % 128;
      sf[(static_cast<size_t>(rb) * num_kb + kb) * 512 + (rr % 32) * 16 + (rr / 32) * 4 + j] = static_cast<uint8_t>(e8);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* pfn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(pfn);
  }
  return fn;
}
static int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_bytes,
                     uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 901;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    cudaFree(nullptr);
    r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? 0 : 900 + static_cast<int>(r);
}

}  // namespace mx
}  // namespace qn

// x [rows, K] bf16 -> q [rows, K] e4m3 + scale atoms sf [ceil(rows / 128) * (K / 128) * 512] (caller zero-fills sf when
// rows % 128 != 0).  K % 128 == 0, ldx % 8 == 0, ldq % 16 == 0.
extern "C" int qn_quantize_mxfp8(const void* x, void* q, void* sf, int rows, int K, long long ldx, long long ldq,
                                 cudaStream_t s) {
  if (rows <= 0) return 0;
  if ((K % 128) || (ldx % 8) || (ldq % 16)) return 930;
  const long long nvec = static_cast<long long>(rows) * (K / 8);
  int grid = static_cast<int>((nvec + 255) / 256 > 148 * 16 ? 148 * 16 : (nvec + 255) / 256);
  qn::mx::quantize_mxfp8_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(q),
                                                     static_cast<uint8_t*>(sf), rows, K, ldx, ldq);
  return static_cast<int>(cudaGetLastError());
}

// C[M, N] (bf16, ldc elements) = (A8 ⊙ SFA) · (B8 ⊙ SFB)ᵀ (+ bias[N]);  a8 [M, K], b8 [N, K] e4m3 (lda / ldb in bytes).
extern "C" int qn_gemm_mxfp8(const void* a8, const void* sfa, const void* b8, const void* sfb, void* c, int M, int N, int K,
                             long long lda, long long ldb, long long ldc, const void* bias, cudaStream_t stream) {
  using namespace qn::mx;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % 128) || (lda % 16) || (ldb % 16) || (ldc % 8) || (bias && (N % 8)) || (reinterpret_cast<uintptr_t>(a8) & 15) ||
      (reinterpret_cast<uintptr_t>(b8) & 15) || (reinterpret_cast<uintptr_t>(c) & 15) ||
      (reinterpret_cast<uintptr_t>(sfa) & 15) || (reinterpret_cast<uintptr_t>(sfb) & 15))
    return 931;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, a8, M, K, lda, BKB, BM))) return rc;
  if ((rc = make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b8, N, K, ldb, BKB, BN))) return rc;
  if ((rc = make_tmap(&tc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, c, M, N, ldc * 2, 64, 32))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const uint32_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  Params p{static_cast<uint32_t>(M), static_cast<uint32_t>(N), static_cast<uint32_t>(K), static_cast<const uint8_t*>(sfa),
           static_cast<const uint8_t*>(sfb), static_cast<const __nv_bfloat16*>(bias)};
  gemm_mxfp8_kernel<<<tiles < static_cast<uint32_t>(sms) ? tiles : sms, kThreads, SMEM_BYTES, stream>>>(ta, tb, tc, p);
  return static_cast<int>(cudaGetLastError());
}
