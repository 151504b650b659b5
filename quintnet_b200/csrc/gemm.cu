// quintnet_b200 — persistent warp-specialised bf16 GEMM on tcgen05 / TMEM / TMA (sm_100a).
//
//   C[M,N] (bf16) = epilogue( A · B ),  fp32 accumulation in tensor memory.
//
// A is the [M x K] operand, B the [N x K] operand; each may be stored
//   * K-major   (row-major [rows=MN, cols=K])   — e.g. activations X[M,K], weights W[N,K]
//   * MN-major  (row-major [rows=K,  cols=MN])  — e.g. dY as A of the wgrad, W as B of the dgrad
// so forward (X·Wᵀ), dgrad (dY·W) and wgrad (dYᵀ·X) all run here without a transpose pass.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0   : TMA producer  — cp.async.bulk.tensor 2-D tiles into a kStages-deep smem ring (128B swizzle)
//   warp 1   : MMA issuer    — one thread issues tcgen05.mma (128 x BLOCK_N x 16), commits to mbarriers
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue     — tcgen05.ld accumulators -> bias / GELU / GELU' -> bf16 -> swizzled smem -> TMA store
// Accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Tiles are walked in a grouped order for L2 reuse.
#include "common.cuh"

#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace qn {

constexpr uint32_t BLOCK_M = 128;
constexpr uint32_t BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle atom
constexpr uint32_t UMMA_K = 16;
constexpr uint32_t kNumThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr uint32_t kEpiWarp0 = 4;
constexpr uint32_t kGroupM = 8;

enum Epilogue : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_MUL_AUX = 3 };

struct GemmParams {
  uint32_t M, N, K;
  const __nv_bfloat16* bias;   // [N] or null
  int epilogue;
  int has_z;                   // EPI_BIAS_GELU: also store gelu'(z) through tmap_z (consumed by EPI_MUL_AUX)
  uint32_t group_m;            // tile rasterisation: m-blocks per L2-reuse group (0 = kGroupM)
  uint32_t splits;             // split-K factor (>1: fp32 partials are atomically added into acc32)
  float* acc32;                // [M, ld32] fp32, zero-initialised by the caller
  uint32_t ld32;
  int acc_f32;                 // the fp32 buffer IS the result (C += A·B in fp32, e.g. main-grad accumulation): every
                               // work unit adds its accumulator with vector red.add, also when splits == 1
  // FP8 (e4m3) operands with row-wise scaling: A and B hold fp8 bytes (K counts byte PAIRS here, so every tile,
  // descriptor and TMA box is byte-identical to the bf16 path); C = (A8 · B8ᵀ) * scale_a[row] * scale_b[col]
  int fp8;
  const float* scale_a;        // [M] dequantisation scale of each A row
  const float* scale_b;        // [N] dequantisation scale of each B row (= output column)
};

// Fused tensor-parallel collectives (FUSE template parameter of the kernel).  Every byte that crosses NVLink is moved
// by the TMA engine from inside the GEMM kernel (1-D bulk async copies on peer-mapped addresses), so no LSU
// instruction, register or per-thread fence of the compute pipeline is spent on the transfer:
//   FUSE_AG (K3): the A operand is the all-gather of every rank's row shard.  Local rows are read straight from
//     x_local.  ONE thread per CTA (warp 3) is a copy pump: it bulk-loads 8-row units of this rank's shard into a
//     small shared-memory ring and bulk-stores each unit into every peer's (and the local) symmetric gather
//     buffer, publishing the unit's flag once its store group has completed.  148 pumps x 2 slots x 16 KB keep
//     ~4.7 MB in flight — far more than the NVLink bandwidth-delay product — while each pump handles only ~7 units.
//     The TMA producer acquires the 16 flags of a remote 128-row block before its first load; tiles are ordered
//     local-rows-first, so the transfer hides behind the local half of the GEMM.
//   FUSE_RS (K4): the output is reduce-scattered by rows.  Partial [32 x 64] slabs whose rows belong to a peer are
//     staged in shared memory (the same swizzled image a TMA tile store would read) and bulk-stored as one 4 KB
//     block into that peer's symmetric staging buffer; their flags are published lazily (two store groups behind)
//     so the epilogue never waits for an NVLink round trip.  Tiles owned by this rank are processed LAST; their
//     epilogue bulk-loads the peers' staged slabs (flag acquire -> 4 KB bulk load) while the accumulator is read
//     from TMEM, adds them in fp32 and stores the tile.  Two staging buffers per epilogue warp.
enum Fuse : int { FUSE_NONE = 0, FUSE_AG = 1, FUSE_RS = 2 };
constexpr int TP_MAX = 8;
enum AgMode : uint32_t { AG_PUMP = 0, AG_LSU = 1, AG_EXTERNAL = 2 };
struct TpComm {
  int rank, world;
  uint32_t seq;                       // launch sequence number of this workspace (flags hold seq values)
  uint32_t rows_per_rank, mb_per_rank;
  const __nv_bfloat16* x_local;       // AG: [rows_per_rank, K] contiguous
  __nv_bfloat16* x_full;              // AG (optional): private [M, K] tensor that receives the COMPLETE gathered operand
                                      // (own rows straight from the pump, peers' rows copied out of the gather buffer
                                      // by the pumps once their flags are up) — the forward keeps it for the wgrad
                                      // GEMM of the backward pass without a separate clone of the gather buffer
  __nv_bfloat16* gather[TP_MAX];      // AG: every rank's gather buffer [M, K] (peer mappings)
  uint32_t* ag_flags[TP_MAX];         // AG: every rank's flags: [kAgDoneBase + src] = "src's whole shard has landed" (pump /
                                      // external), [kLocalCntBase] = this rank's grid-arrival counter; per-8-row flags
                                      // [M / 8] in the round-1 LSU mode
  __nv_bfloat16* stage[TP_MAX];       // RS: every rank's staging buffer [world][slabs_per_rank] x 4 KB slab images
  uint32_t* rs_flags[TP_MAX];         // RS: every rank's flags: [kAgDoneBase + src] = "all of src's partial slabs for this rank
                                      // have landed", [kLocalCntBase] = this rank's epilogue-warp arrival counter
  uint32_t rs_parts;                  // RS: number of separately published row parts per owner (1 or kAgParts)
  uint32_t ag_parts;                  // AG pump: number of separately published shard parts (1 or kAgParts)
  uint32_t ag_mode;                   // AG_PUMP: one bulk-copy pump thread per CTA (default); AG_LSU: two push warps
                                      // per CTA with 16-byte stores (round-1 kernel, kept for A/B timing);
                                      // AG_EXTERNAL: the shard is copied by someone else (copy engine on a side
                                      // stream) who publishes ONE flag per source rank at ag_flags[kAgDoneBase + src]
  uint32_t debug;                     // QN_TP_DEBUG timing-attribution knobs (results are WRONG when set):
                                      //   1 RS: skip the peer wait + add    4 AG: skip the flag waits    8 AG: skip the push
};

constexpr uint32_t kAgDoneBase = 200000;      // flag index of the per-source "whole shard has landed" flags (external push)
constexpr uint32_t kAgParts = 4;              // AG: the shard is published in this many row-contiguous parts (pieces are dealt
                                              // in address order, so part q of every source completes at ~(q+1)/4 of the
                                              // transfer and the consumer can start on it while the rest is in flight)
constexpr uint32_t kLocalCntBase = kAgDoneBase + 64;   // index of this rank's own grid-arrival counters (same flag buffer)
constexpr uint32_t kPumpSlot = 8192;          // AG pump: bytes per ring slot
constexpr uint32_t kPumpRing = 4;             // 4 x 8 KB, loads issued kPumpAhead pieces ahead of the stores
constexpr uint32_t kPumpAhead = 2;
constexpr uint32_t kSmemLimit = 232448;       // 227 KB per CTA on sm_100a

QN_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <uint32_t BLOCK_N, bool CTA2 = false, int FUSE = FUSE_NONE>
struct Cfg {
  // CTA2: a pair of CTAs (cta_group::2) owns a 256 x BLOCK_N tile; each CTA stages its own 128 rows of A and HALF of
  // the B rows, so a pipeline stage is 32 KB instead of 48 KB per SM (a third less L2 -> SM traffic per flop).
  static constexpr uint32_t B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
  static constexpr uint32_t B_BYTES = B_ROWS * BLOCK_K * 2;         // 16/32 KB
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t EPI_PER_WARP = FUSE == FUSE_RS ? 8192 : 4096;   // RS: two (32 rows x 128 B) staging buffers
  static constexpr uint32_t EPI_BYTES = 8 * EPI_PER_WARP;
  static constexpr uint32_t RING_BYTES = FUSE == FUSE_AG ? kPumpRing * kPumpSlot : 0;
  static constexpr uint32_t BAR_BYTES = 512;
  static constexpr uint32_t FIXED_BYTES = EPI_BYTES + RING_BYTES + BAR_BYTES + 1024;   // +align slack
  static constexpr uint32_t kStagesFit = (kSmemLimit - FIXED_BYTES) / STAGE_BYTES;
  static constexpr uint32_t kStages = kStagesFit < 6 ? kStagesFit : 6;      // plain: 4 (256 x 1 CTA) / 6; fused: 3 / 5
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + FIXED_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;                // 256 / 512
  static_assert(kStages >= 3, "pipeline too shallow");
};

// Exit handshake of a CTA pair with a deadline: each CTA arrives on BOTH CTAs' exit barrier and waits on its own
// (mbar_wait carries the cycle-counter watchdog), instead of an untimed barrier.cluster at the end of the kernel —
// a CTA that left early (the round-1 hang) now traps its partner instead of hanging the GPU.
QN_DEVICE void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

QN_DEVICE void tile_coords(uint32_t tile, uint32_t tiles_m, uint32_t tiles_n, uint32_t group_m, uint32_t& m_blk,
                           uint32_t& n_blk) {
  const uint32_t gm = group_m ? group_m : kGroupM;
  const uint32_t per_group = gm * tiles_n;
  const uint32_t group = tile / per_group;
  const uint32_t first_m = group * gm;
  const uint32_t gsize = min(tiles_m - first_m, gm);
  const uint32_t in_group = tile - group * per_group;
  m_blk = first_m + in_group % gsize;
  n_blk = in_group / gsize;
}

template <uint32_t BLOCK_N, bool A_MN, bool B_MN, int FUSE, bool CTA2 = false, bool ACCUM = false, bool FP8 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_z,
                 const __grid_constant__ CUtensorMap tmap_x, const GemmParams p, const TpComm tc) {
  using C = Cfg<BLOCK_N, CTA2, FUSE>;
  constexpr uint32_t kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + kStages * C::STAGE_BYTES;
  uint8_t* smem_ring = smem_epi + C::EPI_BYTES;                       // FUSE_AG: copy-pump ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ring + C::RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_bar = tmem_empty + 2;          // [8] one per epilogue warp (aux / old-C / staged-partial tile loads)
  uint64_t* pump_bar = aux_bar + 8;            // [kPumpRing] FUSE_AG copy pump
  uint64_t* exit_bar = pump_bar + kPumpRing;   // CTA pair: exit handshake
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(exit_bar + 1);
  static_assert((2 * 6 + 4 + 8 + kPumpRing + 1) * 8 + 4 <= C::BAR_BYTES, "barrier area too small");

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;               // 0 = leader of the CTA pair
  const uint32_t n_cta = CTA2 ? gridDim.x / 2 : gridDim.x;               // persistent workers (CTAs or CTA pairs)
  const uint32_t worker = CTA2 ? blockIdx.x / 2 : blockIdx.x;
  constexpr uint32_t TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
  const uint32_t tiles_m = (p.M + TILE_M - 1) / TILE_M;
  const uint32_t tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const uint32_t num_tiles = tiles_m * tiles_n;
  const uint32_t num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const uint32_t splits = p.splits > 1 ? p.splits : 1;
  const uint32_t kb_per_split = (num_kb_total + splits - 1) / splits;
  const uint32_t num_work = num_tiles * splits;
  // tile order: AG walks this rank's own row blocks first, RS walks them last
  constexpr uint32_t kMbPerTile = CTA2 ? 2 : 1;       // 128-row blocks per tile row (tc.mb_per_rank counts 128-row blocks)
  const uint32_t m_shift = FUSE == FUSE_AG ? tc.rank * (tc.mb_per_rank / kMbPerTile)
                         : FUSE == FUSE_RS ? ((tc.rank + 1) % tc.world) * (tc.mb_per_rank / kMbPerTile) : 0u;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    prefetch_tensormap(&tmap_c);
    if (p.has_z || p.epilogue == EPI_MUL_AUX) prefetch_tensormap(&tmap_z);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CTA2 ? 16 : 8);        // CTA pair: the epilogue warps of BOTH CTAs release the leader
    }
    for (uint32_t i = 0; i < 8; ++i) mbar_init(&aux_bar[i], 1);
    for (uint32_t i = 0; i < kPumpRing; ++i) mbar_init(&pump_bar[i], 1);
    mbar_init(exit_bar, 2);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_2cta<C::TMEM_COLS>(tmem_ptr); else tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  }
  tcgen05_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs are live before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    {
      uint32_t stage = 0, phase = 0;
      for (uint32_t work = worker; work < num_work; work += n_cta) {
        const uint32_t tile = work / splits, split = work - tile * splits;
        const uint32_t kb_begin = split * kb_per_split, kb_end = min(num_kb_total, kb_begin + kb_per_split);
        uint32_t m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
        m_blk = (m_blk + m_shift) % tiles_m;
        // CTA pair: this CTA stages rows [rank*128, +128) of the 256-row A tile and B rows [rank*B_ROWS, +B_ROWS)
        const int32_t m0 = m_blk * TILE_M + cta_rank * BLOCK_M, n0 = n_blk * BLOCK_N + cta_rank * (CTA2 ? C::B_ROWS : 0u);
        bool a_remote = false;
        if constexpr (FUSE == FUSE_AG) {
          const uint32_t mb128 = m_blk * kMbPerTile + cta_rank;          // this CTA's 128-row block of A
          a_remote = (mb128 / tc.mb_per_rank) != static_cast<uint32_t>(tc.rank);
          if (a_remote && !(tc.debug & 4u)) {   // the peer's pushes of these 128 rows must have landed
            if (tc.ag_mode == AG_PUMP) {  // one flag per (source rank, shard part): "that part of its shard has landed"
              const uint32_t srcr = mb128 / tc.mb_per_rank;
              const uint32_t part = tc.ag_parts > 1 ? (mb128 % tc.mb_per_rank) * tc.ag_parts / tc.mb_per_rank : 0u;
              if (lane == 0) wait_flag_ge(tc.ag_flags[tc.rank] + kAgDoneBase + srcr * kAgParts + part, tc.seq);
            } else if (tc.ag_mode == AG_EXTERNAL) {   // copy engine: ONE flag per source rank
              if (lane == 0) wait_flag_ge(tc.ag_flags[tc.rank] + kAgDoneBase + mb128 / tc.mb_per_rank, tc.seq);
            } else if (lane < 16) {
              wait_flag_ge(tc.ag_flags[tc.rank] + mb128 * 16 + lane, tc.seq);                // 16 8-row units in parallel
            }
            __syncwarp();
            fence_proxy_async_all();      // generic-proxy (NVLink) writes -> async-proxy (TMA) reads
          }
        }
        // the whole warp walks the K loop (uniform values stay in uniform registers); one elected lane issues the TMA
        for (uint32_t kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const int32_t k0 = kb * BLOCK_K;
          if constexpr (CTA2) {
            // both CTAs load into their own smem; all bytes are counted on the LEADER's full barrier, which the
            // leader arms for the pair (the MMA issuer is the only consumer of it)
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            if constexpr (FUSE == FUSE_AG) {
              if (a_remote) tma_load_2d_2cta(sa, &tmap_x, &full_bar[stage], k0, m0);
              else tma_load_2d_2cta(sa, &tmap_a, &full_bar[stage], k0, m0 - tc.rank * static_cast<int32_t>(tc.rows_per_rank));
            } else if constexpr (!A_MN) {
              tma_load_2d_2cta(sa, &tmap_a, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (uint32_t j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d_2cta(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + 64 * j, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d_2cta(sb, &tmap_b, &full_bar[stage], k0, n0);         // box {64 K, B_ROWS}
            } else {
#pragma unroll
              for (uint32_t j = 0; j < C::B_ROWS / 64; ++j)
                tma_load_2d_2cta(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + 64 * j, k0);
            }
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if constexpr (FUSE == FUSE_AG) {
            if (a_remote) tma_load_2d(sa, &tmap_x, &full_bar[stage], k0, m0);
            else tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0 - tc.rank * static_cast<int32_t>(tc.rows_per_rank));
          } else if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);                 // box {64 K, 128 M}
          } else {
#pragma unroll
            for (uint32_t j = 0; j < BLOCK_M / 64; ++j)                           // box {64 M, 64 K} slabs
              tma_load_2d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + 64 * j, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);                 // box {64 K, BLOCK_N}
          } else {
#pragma unroll
            for (uint32_t j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + 64 * j, k0);
          }
          }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    // The WHOLE warp walks the loop (all values warp-uniform -> uniform registers); only the tcgen05 instructions are
    // predicated on one elected lane.  Issuing from inside a divergent `lane == 0` region instead made the compiler
    // wrap every MMA in an ELECT / R2UR / branch "waterfall" (~40 SASS instructions per MMA on a single thread): the
    // issue loop, not the tensor pipe, paced the kernel (ncu: tensor pipe 57 % active).  Descriptors are a per-stage
    // base plus a constant per-k increment of the 14-bit address field.
    if (cta_rank == 0) {                              // CTA pair: only the leader issues (for both SMs); warp-uniform
      constexpr uint32_t idesc = FP8 ? make_idesc_e4m3(TILE_M, BLOCK_N)
                                     : make_idesc_bf16(TILE_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      // K-major: 8-row groups are 1024 B apart, one atom along K.  MN-major: 8-K-row groups 1024 B apart,
      // 64-wide MN atoms BLOCK_K*128 B apart.
      constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 0, B_LBO = B_MN ? BLOCK_K * 128 : 0;
      constexpr uint32_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;      // in descriptor address units (16 B)
      constexpr uint32_t B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint64_t a_hi = make_smem_desc_sw128(0, 1024, A_LBO), b_hi = make_smem_desc_sw128(0, 1024, B_LBO);
      const uint32_t smem_base = smem_u32(smem);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (uint32_t work = worker; work < num_work; work += n_cta) {
        const uint32_t split = work % splits;
        const uint32_t kb_begin = split * kb_per_split, kb_end = min(num_kb_total, kb_begin + kb_per_split);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (uint32_t kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t adesc = a_hi | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
          const uint64_t bdesc = b_hi | static_cast<uint64_t>(((sa + C::A_BYTES) & 0x3FFFF) >> 4);
          const uint32_t acc0 = kb != kb_begin ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint32_t acc = k ? 1u : acc0;
              if constexpr (CTA2) {
                if constexpr (FP8) umma_f8_ss_2cta(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, acc);
                else umma_bf16_ss_2cta(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, acc);
              } else {
                if constexpr (FP8) umma_f8_ss(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, acc);
                else umma_bf16_ss(tmem_d, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, acc);
              }
            }
            // frees the smem slot when these MMAs retire (CTA pair: in both CTAs, each has its own producer)
            if constexpr (CTA2) umma_commit_2cta(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            // accumulator complete -> epilogue (CTA pair: the epilogues of both CTAs)
            if (kb + 1 == kb_end) { if constexpr (CTA2) umma_commit_2cta(&tmem_full[as]); else umma_commit(&tmem_full[as]); }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (FUSE == FUSE_AG && warp == 3 && tc.ag_mode == AG_PUMP) {
    // ------------------------------------------------ all-gather copy pump (this rank's shard -> every rank)
    // One thread per CTA.  The shard (contiguous in x_local and in every gather buffer) is cut into kPumpSlot-byte
    // pieces dealt round-robin over the CTAs; a piece moves bulk load -> ring slot -> one bulk store per destination
    // (peers over NVLink + the local copy, which makes the complete gathered operand available to the weight-gradient
    // GEMM), loads running kPumpAhead pieces ahead of the stores.  NO per-piece completion wait or fence: a system
    // fence drains the SM's outstanding NVLink writes, so publishing per unit serialised the pump on the link round
    // trip (measured: round-1 LSU push and the first TMA pump both cost ~25-30 us per kernel).  Instead every pump
    // waits ONCE for its store groups, fences once and arrives on a local counter; the last arrival of the grid
    // publishes ONE flag per peer: "rank r's whole shard has landed".
    if (lane == 0) {
      const size_t shard_bytes = static_cast<size_t>(tc.rows_per_rank) * p.K * 2;
      const uint32_t npieces = static_cast<uint32_t>((shard_bytes + kPumpSlot - 1) / kPumpSlot);
      const uint32_t mine = (tc.debug & 8u) ? 0u : (npieces > blockIdx.x ? (npieces - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(tc.x_local);
      const size_t shard_off = static_cast<size_t>(tc.rank) * shard_bytes;
      uint8_t* local_dst = reinterpret_cast<uint8_t*>(tc.x_full ? tc.x_full : tc.gather[tc.rank]);
      uint32_t cnt = 0;                              // ring uses so far (slot = cnt % kPumpRing)
      auto run = [&](const uint8_t* from, size_t base_off, bool to_peers, uint32_t n_mine, size_t total_bytes) {
        for (uint32_t it = 0; it < n_mine + kPumpAhead; ++it) {
          if (it < n_mine) {
            const size_t off = static_cast<size_t>(blockIdx.x + it * gridDim.x) * kPumpSlot;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<size_t>(kPumpSlot), total_bytes - off));
            const uint32_t slot = (cnt + it) % kPumpRing;
            tma_store_wait_read<kPumpRing - kPumpAhead - 1>();     // the store that last read this slot has drained it
            mbar_arrive_expect_tx(&pump_bar[slot], bytes);
            bulk_load_1d(smem_ring + slot * kPumpSlot, from + off, bytes, &pump_bar[slot]);
          }
          if (it >= kPumpAhead) {
            const uint32_t j = it - kPumpAhead;
            const size_t off = static_cast<size_t>(blockIdx.x + j * gridDim.x) * kPumpSlot;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<size_t>(kPumpSlot), total_bytes - off));
            const uint32_t slot = (cnt + j) % kPumpRing;
            uint8_t* buf = smem_ring + slot * kPumpSlot;
            mbar_wait(&pump_bar[slot], ((cnt + j) / kPumpRing) & 1);
            if (to_peers)
              for (int pr = 1; pr < tc.world; ++pr)
                bulk_store_1d(reinterpret_cast<uint8_t*>(tc.gather[(tc.rank + pr) % tc.world]) + base_off + off, buf, bytes);
            bulk_store_1d(local_dst + base_off + off, buf, bytes);
            tma_store_commit();
          }
        }
        cnt += n_mine;
      };
      // part boundaries: when this CTA has stored its last piece of part q, it waits for its store groups, fences and
      // arrives on the part's local counter; the last arrival of the grid publishes "part q of rank r has landed"
      {
        const uint32_t parts = tc.ag_parts;
        const size_t part_bytes = shard_bytes / parts;             // (multiple of 128 rows: checked on the host)
        uint32_t* lcnt = tc.ag_flags[tc.rank] + kLocalCntBase;
        uint32_t cur = 0;                                          // first part not yet signalled by this CTA
        auto arrive_upto = [&](uint32_t upto) {                    // signal parts [cur, upto)
          if (upto <= cur) return;
          tma_store_wait_all<0>();                                 // every store this CTA issued so far has been performed
          fence_proxy_async_all();
          __threadfence_system();
          for (uint32_t q = cur; q < upto; ++q) {
            uint32_t prev;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(lcnt + q) : "memory");
            if (prev == gridDim.x - 1) {                           // last pump of the grid for this part
              asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lcnt + q), "r"(0u) : "memory");
              __threadfence_system();
              for (int pr = 1; pr < tc.world; ++pr)
                st_relaxed_sys(tc.ag_flags[(tc.rank + pr) % tc.world] + kAgDoneBase + tc.rank * kAgParts + q, tc.seq);
            }
          }
          cur = upto;
        };
        // same pipelined loop as run(), with the part check after each store
        for (uint32_t it = 0; it < mine + kPumpAhead; ++it) {
          if (it < mine) {
            const size_t off = static_cast<size_t>(blockIdx.x + it * gridDim.x) * kPumpSlot;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<size_t>(kPumpSlot), shard_bytes - off));
            const uint32_t slot = (cnt + it) % kPumpRing;
            tma_store_wait_read<kPumpRing - kPumpAhead - 1>();
            mbar_arrive_expect_tx(&pump_bar[slot], bytes);
            bulk_load_1d(smem_ring + slot * kPumpSlot, src + off, bytes, &pump_bar[slot]);
          }
          if (it >= kPumpAhead) {
            const uint32_t j = it - kPumpAhead;
            const size_t off = static_cast<size_t>(blockIdx.x + j * gridDim.x) * kPumpSlot;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<size_t>(kPumpSlot), shard_bytes - off));
            const uint32_t slot = (cnt + j) % kPumpRing;
            uint8_t* buf = smem_ring + slot * kPumpSlot;
            mbar_wait(&pump_bar[slot], ((cnt + j) / kPumpRing) & 1);
            for (int pr = 1; pr < tc.world; ++pr)
              bulk_store_1d(reinterpret_cast<uint8_t*>(tc.gather[(tc.rank + pr) % tc.world]) + shard_off + off, buf, bytes);
            bulk_store_1d(local_dst + shard_off + off, buf, bytes);
            tma_store_commit();
            // the next piece of this CTA (if any) starts at off + gridDim.x * kPumpSlot: parts below it are complete here
            const size_t next_off = off + static_cast<size_t>(gridDim.x) * kPumpSlot;
            if (parts > 1) arrive_upto(next_off >= shard_bytes ? parts : static_cast<uint32_t>(next_off / part_bytes));
          }
        }
        cnt += mine;
        arrive_upto(parts);                                        // (CTAs without pieces, or parts == 1)
      }
      // second duty (background): peers' rows, gather buffer -> x_full, once their shards have landed
      if (tc.x_full != nullptr && !(tc.debug & 4u)) {
        for (int pr = 1; pr < tc.world; ++pr) {
          const uint32_t srcr = (tc.rank + pr) % tc.world;
          for (uint32_t q = 0; q < tc.ag_parts; ++q)
            wait_flag_ge(tc.ag_flags[tc.rank] + kAgDoneBase + srcr * kAgParts + q, tc.seq);
          fence_proxy_async_all();
          const uint32_t m2 = npieces > blockIdx.x ? (npieces - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
          run(reinterpret_cast<const uint8_t*>(tc.gather[tc.rank]) + static_cast<size_t>(srcr) * shard_bytes,
              static_cast<size_t>(srcr) * shard_bytes, false, m2, shard_bytes);
        }
        tma_store_wait_all<0>();
      }
    }
  } else if (FUSE == FUSE_AG && tc.ag_mode == AG_LSU && (warp == 2 || warp == 3)) {
    // ------------------------------------------------ round-1 all-gather push with 16-byte LSU stores (A/B timing only)
    const uint32_t units = tc.rows_per_rank / 32;
    const uint32_t vec_per_unit = 32 * p.K / 8;                 // 16-byte vectors per 32-row unit
    for (uint32_t u = blockIdx.x * 2 + (warp - 2); u < units && !(tc.debug & 8u); u += gridDim.x * 2) {
      const uint4* src = reinterpret_cast<const uint4*>(tc.x_local + static_cast<size_t>(u) * 32 * p.K);
      const size_t dst_off = (static_cast<size_t>(tc.rank) * tc.rows_per_rank + u * 32) * p.K;
      for (uint32_t i = lane; i < vec_per_unit; i += 32 * 16) {
        uint4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) if (i + j * 32 < vec_per_unit) v[j] = ld_nc_v4(src + i + j * 32);
        for (int pr = 0; pr < tc.world; ++pr) {
          uint4* dst = reinterpret_cast<uint4*>(tc.gather[(tc.rank + pr) % tc.world] + dst_off);
#pragma unroll
          for (int j = 0; j < 16; ++j) if (i + j * 32 < vec_per_unit) st_v4(dst + i + j * 32, v[j]);
        }
      }
      __threadfence_system();
      __syncwarp();
      if (lane < 4)                                        // flags are per 8 rows: this unit covers four
        for (int pr = 1; pr < tc.world; ++pr)
          st_relaxed_sys(tc.ag_flags[(tc.rank + pr) % tc.world] + (tc.rank * units + u) * 4 + lane, tc.seq);
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------------------------------------------------------- epilogue (8 warps)
    // warp -> TMEM lane quarter (warp % 4) and column half ((warp - 4) / 4): two warps share each 32-row
    // slab and split its 64-column chunks, doubling the elementwise throughput behind the MMA pipe.
    const uint32_t q = (warp - kEpiWarp0) & 3;
    const uint32_t half = (warp - kEpiWarp0) >> 2;
    constexpr uint32_t kChunks = BLOCK_N / 64;
    uint8_t* sbuf = smem_epi + (warp - kEpiWarp0) * C::EPI_PER_WARP;
    uint8_t* srow = sbuf + lane * 128;
    uint32_t as = 0, aphase = 0, aux_phase = 0;
    // FUSE_RS: slab counter (staging buffer = slab & 1; every slab commits exactly one bulk group).  Remote slabs are
    // fire-and-forget: no per-slab completion wait, fence or flag (a system fence drains the SM's outstanding NVLink
    // writes — one per slab put the link round trip on the epilogue's critical path).  When a warp has issued its last
    // remote slab (i.e. at its first own slab, or at the end of its tile loop) it waits ONCE for its store groups,
    // fences once and arrives on a local counter; the last of the grid's epilogue warps publishes ONE flag per peer:
    // "every partial slab rank r owes you has landed".  Own tiles come last, so that flag is normally up already.
    uint32_t rs_slab = 0;
    // Publication is per (owner, row part): the tile order walks the owners' row ranges in sequence (peers first, own rows
    // last; part boundaries coincide with tile-group boundaries), so "key" = owner-order * parts + part is monotone along
    // the loop.  A warp arrives on key k's counter when it moves past it; the grid's last arrival tells that owner
    // "part q of rank r's partials has landed" — the owner starts reducing its first rows while the peers are still
    // producing the later parts.
    uint32_t rs_cur = 0, rs_ready = 0;         // first key not yet signalled / first own part not yet confirmed
    const uint32_t rs_keys = (tc.world - 1) * tc.rs_parts;
    auto rs_arrive_upto = [&](uint32_t upto) { // lane 0 only
      if (upto <= rs_cur) return;
      tma_store_wait_all<0>();
      fence_proxy_async_all();
      __threadfence_system();
      uint32_t* lcnt = tc.rs_flags[tc.rank] + kLocalCntBase;
      for (uint32_t k = rs_cur; k < upto; ++k) {
        uint32_t prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(lcnt + k) : "memory");
        if (prev == gridDim.x * 8 - 1) {       // last epilogue warp of the grid for this key
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lcnt + k), "r"(0u) : "memory");
          __threadfence_system();
          const uint32_t dst = (tc.rank + 1 + k / tc.rs_parts) % tc.world;
          st_relaxed_sys(tc.rs_flags[dst] + kAgDoneBase + tc.rank * kAgParts + k % tc.rs_parts, tc.seq);
        }
      }
      rs_cur = upto;
    };
    for (uint32_t work = worker; work < num_work; work += n_cta) {
      const uint32_t tile = work / splits;
      uint32_t m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
      m_blk = (m_blk + m_shift) % tiles_m;
      const uint32_t m0 = m_blk * TILE_M + cta_rank * BLOCK_M + q * 32, n0 = n_blk * BLOCK_N;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
#pragma unroll 1
      for (uint32_t ci = 0; ci < kChunks / 2; ++ci) {
        const uint32_t chunk = half * (kChunks / 2) + ci;
        const uint32_t nc = n0 + chunk * 64;
        if (nc >= p.N) break;                          // warp-uniform
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BLOCK_N + chunk * 64;
        // the TMA store that last read this staging buffer must have finished reading it (FUSE_RS double-buffers
        // its staging and waits inside its own branch)
        if constexpr (FUSE != FUSE_RS) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if (p.splits > 1 || p.acc_f32) {
          // split-K / fp32 accumulation: fp32 partial of this K-slice -> vectorised atomics into the fp32 buffer
          const uint32_t row = m0 + lane;
#pragma unroll
          for (uint32_t hcol = 0; hcol < 2; ++hcol) {
            uint32_t acc[32];
            tmem_ld_32x32b_x32(taddr + hcol * 32, acc);
            tmem_ld_wait();
            if (row < p.M) {
              float* dst = p.acc32 + static_cast<size_t>(row) * p.ld32 + nc + hcol * 32;
#pragma unroll
              for (uint32_t v = 0; v < 8; ++v) {
                if (nc + hcol * 32 + v * 4 < p.N)
                  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dst + v * 4),
                               "f"(__uint_as_float(acc[v * 4 + 0])), "f"(__uint_as_float(acc[v * 4 + 1])),
                               "f"(__uint_as_float(acc[v * 4 + 2])), "f"(__uint_as_float(acc[v * 4 + 3]))
                               : "memory");
              }
            }
          }
          continue;
        }
        if constexpr (FUSE == FUSE_RS) {
          const uint32_t mb128 = m_blk * kMbPerTile + cta_rank;
          const uint32_t owner = mb128 / tc.mb_per_rank;
          const uint32_t m_loc = m0 - owner * tc.rows_per_rank;          // first row of this warp's slab in the owner's shard
          const uint32_t n_chunks = (p.N + 63) / 64;
          const uint32_t slab_id = ((mb128 % tc.mb_per_rank) * 4 + q) * n_chunks + (nc >> 6);
          const uint32_t slabs_per_rank = tc.mb_per_rank * 4 * n_chunks;
          const bool mine = owner == static_cast<uint32_t>(tc.rank);
          uint8_t* obuf = sbuf + (rs_slab & 1u) * 4096;                  // output staging of this slab
          uint8_t* pbuf = sbuf + ((rs_slab & 1u) ^ 1u) * 4096;           // landing zone of the peers' staged partials
          uint8_t* orow = obuf + lane * 128;
          uint64_t* ab = &aux_bar[warp - kEpiWarp0];
          ++rs_slab;
          if (lane == 0) {
            const uint32_t part = (mb128 % tc.mb_per_rank) * tc.rs_parts / tc.mb_per_rank;
            // everything this warp owed for the keys before this slab's is out
            rs_arrive_upto(mine ? rs_keys : ((owner + tc.world - tc.rank - 1) % tc.world) * tc.rs_parts + part);
            if (mine && !(tc.debug & 1u)) {
              tma_store_wait_read<0>();            // both staging buffers are free
              const uint32_t src = (tc.rank + 1) % tc.world;       // first peer's partial: load overlaps the TMEM read
              for (; rs_ready <= part; ++rs_ready)     // every peer's partial slabs for this part of my rows have landed
                for (int pr = 1; pr < tc.world; ++pr)
                  wait_flag_ge(tc.rs_flags[tc.rank] + kAgDoneBase + ((tc.rank + pr) % tc.world) * kAgParts + rs_ready, tc.seq);
              fence_proxy_async_all();
              mbar_arrive_expect_tx(ab, 4096);
              bulk_load_1d(pbuf, reinterpret_cast<const uint8_t*>(tc.stage[tc.rank]) +
                                     (static_cast<size_t>(src) * slabs_per_rank + slab_id) * 4096, 4096, ab);
            } else {
              tma_store_wait_read<1>();            // the group that last read obuf (two slabs ago) has drained it
            }
          }
          __syncwarp();
          uint32_t acc[64];
          tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
          tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          tmem_ld_wait();
          float* f = reinterpret_cast<float*>(acc);
          if (p.bias != nullptr) {       // the bias lives on ONE rank (tp-rank 0) and rides on every partial slab of that rank
#pragma unroll
            for (uint32_t v = 0; v < 8; ++v) {
              if (nc + v * 8 < p.N) {
                const uint4 b = ld_nc_v4(p.bias + nc + v * 8);
                const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
                f[v * 8 + 0] += b0.x; f[v * 8 + 1] += b0.y; f[v * 8 + 2] += b1.x; f[v * 8 + 3] += b1.y;
                f[v * 8 + 4] += b2.x; f[v * 8 + 5] += b2.y; f[v * 8 + 6] += b3.x; f[v * 8 + 7] += b3.y;
              }
            }
          }
          if (mine && !(tc.debug & 1u)) {
            // reduce: add every peer's staged partial of this [32 x 64] slab (same swizzled image the sender staged)
            const uint8_t* prow = pbuf + lane * 128;
            for (int pr = 1; pr < tc.world; ++pr) {
              mbar_wait(ab, aux_phase);
              aux_phase ^= 1;
#pragma unroll
              for (uint32_t v = 0; v < 8; ++v) {
                const uint4 a4 = *reinterpret_cast<const uint4*>(prow + ((v ^ (lane & 7)) << 4));
                const float2 a0 = unpack_bf16(a4.x), a1 = unpack_bf16(a4.y), a2 = unpack_bf16(a4.z), a3 = unpack_bf16(a4.w);
                f[v * 8 + 0] += a0.x; f[v * 8 + 1] += a0.y; f[v * 8 + 2] += a1.x; f[v * 8 + 3] += a1.y;
                f[v * 8 + 4] += a2.x; f[v * 8 + 5] += a2.y; f[v * 8 + 6] += a3.x; f[v * 8 + 7] += a3.y;
              }
              __syncwarp();
              if (pr + 1 < tc.world && lane == 0) {                 // next peer's partial into the same landing zone
                const uint32_t src = (tc.rank + pr + 1) % tc.world;
                mbar_arrive_expect_tx(ab, 4096);
                bulk_load_1d(pbuf, reinterpret_cast<const uint8_t*>(tc.stage[tc.rank]) +
                                       (static_cast<size_t>(src) * slabs_per_rank + slab_id) * 4096, 4096, ab);
              }
            }
          }
#pragma unroll
          for (uint32_t v = 0; v < 8; ++v) {
            uint4 pk;
            pk.x = pack_bf16(f[v * 8 + 0], f[v * 8 + 1]); pk.y = pack_bf16(f[v * 8 + 2], f[v * 8 + 3]);
            pk.z = pack_bf16(f[v * 8 + 4], f[v * 8 + 5]); pk.w = pack_bf16(f[v * 8 + 6], f[v * 8 + 7]);
            *reinterpret_cast<uint4*>(orow + ((v ^ (lane & 7)) << 4)) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (mine) {
              tma_store_2d(&tmap_c, obuf, static_cast<int32_t>(nc), static_cast<int32_t>(m_loc));
              tma_store_commit();
            } else {
              // scatter: ONE 4 KB bulk store of the slab image into the owner's staging buffer over NVLink
              bulk_store_1d(reinterpret_cast<uint8_t*>(tc.stage[owner]) +
                                (static_cast<size_t>(tc.rank) * slabs_per_rank + slab_id) * 4096, obuf, 4096);
              tma_store_commit();
            }
          }
          continue;
        }
        if constexpr (ACCUM) {
          // C += A·B (weight-gradient accumulation over micro-batches straight into the gradient buffer): the old
          // C tile is bulk-loaded by TMA into the staging buffer, the accumulator is added and the tile stored back
          uint64_t* ab = &aux_bar[warp - kEpiWarp0];
          if (lane == 0) {
            mbar_arrive_expect_tx(ab, 4096);
            tma_load_2d(sbuf, &tmap_c, ab, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
          }
          uint32_t acc[64];
          tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
          tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          tmem_ld_wait();
          mbar_wait(ab, aux_phase);
          aux_phase ^= 1;
          const float* f = reinterpret_cast<const float*>(acc);
#pragma unroll
          for (uint32_t v = 0; v < 8; ++v) {
            uint4* sp = reinterpret_cast<uint4*>(srow + ((v ^ (lane & 7)) << 4));
            const uint4 a4 = *sp;
            const float2 a0 = unpack_bf16(a4.x), a1 = unpack_bf16(a4.y), a2 = unpack_bf16(a4.z), a3 = unpack_bf16(a4.w);
            uint4 pk;
            pk.x = pack_bf16(f[v * 8 + 0] + a0.x, f[v * 8 + 1] + a0.y); pk.y = pack_bf16(f[v * 8 + 2] + a1.x, f[v * 8 + 3] + a1.y);
            pk.z = pack_bf16(f[v * 8 + 4] + a2.x, f[v * 8 + 5] + a2.y); pk.w = pack_bf16(f[v * 8 + 6] + a3.x, f[v * 8 + 7] + a3.y);
            *sp = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_c, sbuf, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
            tma_store_commit();
          }
          continue;
        }
        if (p.epilogue == EPI_MUL_AUX) {
          // out = acc ⊙ aux : the aux tile (e.g. gelu'(z)) is bulk-loaded by TMA into the staging buffer,
          // multiplied in place and stored back by TMA — no strided per-thread global loads
          uint64_t* ab = &aux_bar[warp - kEpiWarp0];
          if (lane == 0) {
            mbar_arrive_expect_tx(ab, 4096);
            tma_load_2d(sbuf, &tmap_z, ab, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
          }
          uint32_t acc[64];
          tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
          tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          tmem_ld_wait();
          mbar_wait(ab, aux_phase);
          aux_phase ^= 1;
          const float* f = reinterpret_cast<const float*>(acc);
#pragma unroll
          for (uint32_t v = 0; v < 8; ++v) {
            uint4* sp = reinterpret_cast<uint4*>(srow + ((v ^ (lane & 7)) << 4));
            const uint4 a4 = *sp;
            const float2 a0 = unpack_bf16(a4.x), a1 = unpack_bf16(a4.y), a2 = unpack_bf16(a4.z), a3 = unpack_bf16(a4.w);
            uint4 pk;
            pk.x = pack_bf16(f[v * 8 + 0] * a0.x, f[v * 8 + 1] * a0.y); pk.y = pack_bf16(f[v * 8 + 2] * a1.x, f[v * 8 + 3] * a1.y);
            pk.z = pack_bf16(f[v * 8 + 4] * a2.x, f[v * 8 + 5] * a2.y); pk.w = pack_bf16(f[v * 8 + 6] * a3.x, f[v * 8 + 7] * a3.y);
            *sp = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_c, sbuf, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
            tma_store_commit();
          }
          continue;
        }
        const bool dual = (p.epilogue == EPI_BIAS_GELU);
        uint32_t gpk[32];                                 // gelu'(z) packed bf16 (second store), dual only
#pragma unroll
        for (uint32_t hcol = 0; hcol < 2; ++hcol) {        // two 32-column halves keep registers low
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + hcol * 32, acc);
          tmem_ld_wait();
          float* f = reinterpret_cast<float*>(acc);
          const uint32_t nh = nc + hcol * 32;
          if constexpr (FP8) {                             // dequantise: per-row scale of A x per-column scale of B
            const float sa = (m0 + lane < p.M) ? p.scale_a[m0 + lane] : 0.f;
#pragma unroll
            for (uint32_t v = 0; v < 8; ++v) {
              if (nh + v * 4 < p.N) {
                const float4 sb = *reinterpret_cast<const float4*>(p.scale_b + nh + v * 4);
                f[v * 4 + 0] *= sa * sb.x; f[v * 4 + 1] *= sa * sb.y; f[v * 4 + 2] *= sa * sb.z; f[v * 4 + 3] *= sa * sb.w;
              }
            }
          }
          if (p.epilogue != EPI_NONE && p.bias != nullptr) {
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
            float x[8];
            if (dual) {
              float g[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) gelu_both(f[v * 8 + e], x[e], g[e]);
              gpk[(hcol * 4 + v) * 4 + 0] = pack_bf16(g[0], g[1]); gpk[(hcol * 4 + v) * 4 + 1] = pack_bf16(g[2], g[3]);
              gpk[(hcol * 4 + v) * 4 + 2] = pack_bf16(g[4], g[5]); gpk[(hcol * 4 + v) * 4 + 3] = pack_bf16(g[6], g[7]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = f[v * 8 + e];
            }
            uint4 pk;
            pk.x = pack_bf16(x[0], x[1]); pk.y = pack_bf16(x[2], x[3]);
            pk.z = pack_bf16(x[4], x[5]); pk.w = pack_bf16(x[6], x[7]);
            *reinterpret_cast<uint4*>(srow + (((hcol * 4 + v) ^ (lane & 7)) << 4)) = pk;   // 128B swizzle
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_c, sbuf, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
          tma_store_commit();
        }
        if (dual && p.has_z) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (uint32_t v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(srow + ((v ^ (lane & 7)) << 4)) =
                make_uint4(gpk[v * 4 + 0], gpk[v * 4 + 1], gpk[v * 4 + 2], gpk[v * 4 + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_z, sbuf, static_cast<int32_t>(nc), static_cast<int32_t>(m0));
            tma_store_commit();
          }
        }
      }
      // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if constexpr (FUSE == FUSE_RS) { if (lane == 0) rs_arrive_upto(rs_keys); }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CTA2) {     // the partner may still read this CTA's smem / signal its barriers: leave together (timed)
    if (threadIdx.x == 0) {
      mbar_arrive_cluster(exit_bar, 0);
      mbar_arrive_cluster(exit_bar, 1);
      mbar_wait(exit_bar, 0);
    }
    __syncthreads();
  }
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base); else tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box {box_cols, box_rows}
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                     uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 901;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // first CUDA call on this host thread (e.g. an autograd worker): bind the primary context and retry
    cudaFree(nullptr);
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? 0 : 900 + static_cast<int>(r);
}

static int g_num_sms = 0;
constexpr int kDefault2Cta = 1;           // QN_GEMM_2CTA overrides: 0 off, 1 auto, 2 every BLOCK_N = 256 launch

// CTA pairs or single CTAs?  A pair finishes a 256 x 256 tile ~1.25x faster than two CTAs finish two 128 x 256
// tiles (a third less L2 -> SM traffic), so pairs win unless wave quantisation on 74 pairs is much worse than on
// 148 CTAs.
static bool use_cta_pairs(int M, int N, int block_n, int splits, int max_ctas) {
  static const int mode = [] { const char* e = getenv("QN_GEMM_2CTA"); return e ? atoi(e) : kDefault2Cta; }();
  if (!mode || block_n != 256 || max_ctas != 0) return false;
  if (mode >= 2) return true;
  if (M < 256 || N < 256) return false;
  const long long w2 = static_cast<long long>((M + 255) / 256) * ((N + 255) / 256) * splits;
  const long long w1 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256) * splits;
  const double eff2 = static_cast<double>(w2) / (74.0 * ((w2 + 73) / 74));
  const double eff1 = static_cast<double>(w1) / (148.0 * ((w1 + 147) / 148));
  return eff2 * 1.2 >= eff1;
}

// CTA-pair (cta_group::2) launch: clusters of 2 CTAs, one 256 x 256 tile per cluster and step
template <bool A_MN, bool B_MN, int FUSE = FUSE_NONE, bool ACCUM = false, bool FP8 = false>
static int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tz,
                       const GemmParams& p, cudaStream_t stream, const CUtensorMap* tx = nullptr,
                       const TpComm* comm = nullptr, int max_ctas = 0) {
  using C = Cfg<256, true, FUSE>;
  auto kern = gemm_bf16_kernel<256, A_MN, B_MN, FUSE, true, ACCUM, FP8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const uint32_t tiles = ((p.M + 255) / 256) * ((p.N + 255) / 256) * (p.splits > 1 ? p.splits : 1);
  uint32_t pairs = static_cast<uint32_t>(g_num_sms / 2);
  if (max_ctas > 1 && static_cast<uint32_t>(max_ctas / 2) < pairs) pairs = static_cast<uint32_t>(max_ctas / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs), 1, 1);
  cfg.blockDim = dim3(kNumThreads, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TpComm c0{};
  if (FUSE != FUSE_NONE && comm == nullptr) return 914;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tz, tx ? *tx : tc, p, comm ? *comm : c0);
  return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
}

template <uint32_t BLOCK_N, bool A_MN, bool B_MN, int FUSE = FUSE_NONE, bool ACCUM = false, bool FP8 = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tz,
                  const GemmParams& p, int max_ctas, cudaStream_t stream, const CUtensorMap* tx = nullptr,
                  const TpComm* comm = nullptr) {
  using C = Cfg<BLOCK_N, false, FUSE>;
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, FUSE, false, ACCUM, FP8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const uint32_t tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N) * (p.splits > 1 ? p.splits : 1);
  int grid = static_cast<int>(tiles < static_cast<uint32_t>(g_num_sms) ? tiles : g_num_sms);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  TpComm c0{};
  if (FUSE != FUSE_NONE && comm == nullptr) return 914;
  kern<<<grid, kNumThreads, C::SMEM_BYTES, stream>>>(ta, tb, tc, tz, tx ? *tx : tc, p, comm ? *comm : c0);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace qn

// C ABI -------------------------------------------------------------------------------------------------
// a: A operand.  a_mn=0: row-major [M, K] (lda >= K);  a_mn=1: row-major [K, M] (lda >= M)
// b: B operand.  b_mn=0: row-major [N, K] (ldb >= K);  b_mn=1: row-major [K, N] (ldb >= N)
// c: row-major [M, N] bf16 (ldc >= N).  All leading dimensions must be multiples of 8 elements and all
// base pointers 16-byte aligned.  block_n: 0 = auto, 128 or 256.  max_ctas: 0 = all SMs.
extern "C" int qn_gemm_bf16(const void* a, const void* b, void* c, int M, int N, int K, long long lda,
                            long long ldb, long long ldc, int a_mn, int b_mn, int epilogue, const void* bias,
                            void* z_or_aux, long long ldz, int block_n, int max_ctas, cudaStream_t stream) {
  using namespace qn;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (reinterpret_cast<uintptr_t>(a) & 15) ||
      (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(c) & 15))
    return 910;
  if (block_n == 0) {
    const long long t256 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256);
    block_n = (N > 128 && t256 >= 96) ? 256 : 128;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.epilogue = epilogue;
  p.has_z = (epilogue == EPI_BIAS_GELU && z_or_aux != nullptr) ? 1 : 0;
  if (epilogue == EPI_MUL_AUX && (z_or_aux == nullptr || (ldz % 8))) return 911;
  if ((epilogue == EPI_BIAS || epilogue == EPI_BIAS_GELU) && bias != nullptr && (N % 8)) return 912;

  CUtensorMap ta, tb, tc, tz;
  int rc;
  const bool two_cta = use_cta_pairs(M, N, block_n, 1, max_ctas);
  if (!a_mn) rc = make_tmap(&ta, a, M, K, lda, 64, BLOCK_M); else rc = make_tmap(&ta, a, K, M, lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&tb, b, N, K, ldb, 64, two_cta ? 128 : block_n); else rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K);
  if (rc) return rc;
  rc = make_tmap(&tc, c, M, N, ldc, 64, 32);
  if (rc) return rc;
  if (p.has_z || epilogue == EPI_MUL_AUX) {
    if ((ldz % 8) || (reinterpret_cast<uintptr_t>(z_or_aux) & 15)) return 913;
    rc = make_tmap(&tz, z_or_aux, M, N, ldz, 64, 32);
    if (rc) return rc;
  } else {
    tz = tc;
  }
#define QN_DISPATCH(BN)                                                                                   \
  do {                                                                                                    \
    if (!a_mn && !b_mn) return launch<BN, false, false>(ta, tb, tc, tz, p, max_ctas, stream);             \
    if (!a_mn && b_mn) return launch<BN, false, true>(ta, tb, tc, tz, p, max_ctas, stream);               \
    if (a_mn && !b_mn) return launch<BN, true, false>(ta, tb, tc, tz, p, max_ctas, stream);               \
    return launch<BN, true, true>(ta, tb, tc, tz, p, max_ctas, stream);                                   \
  } while (0)
  if (two_cta) {
    if (!a_mn && !b_mn) return launch_2cta<false, false>(ta, tb, tc, tz, p, stream);
    if (!a_mn && b_mn) return launch_2cta<false, true>(ta, tb, tc, tz, p, stream);
    if (a_mn && !b_mn) return launch_2cta<true, false>(ta, tb, tc, tz, p, stream);
    return launch_2cta<true, true>(ta, tb, tc, tz, p, stream);
  }
  if (block_n == 256) QN_DISPATCH(256);
  QN_DISPATCH(128);
#undef QN_DISPATCH
}


// ------------------------------------------------------------------------------------------- split-K
namespace qn {
__global__ void __launch_bounds__(256)
f32_to_bf16_2d_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols, long long ld_in,
                      long long ld_out) {
  const long long n4 = static_cast<long long>(rows) * (cols / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / (cols / 4)), c = static_cast<int>(i % (cols / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + r * ld_in + c);
    *reinterpret_cast<uint2*>(out + r * ld_out + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}
__global__ void __launch_bounds__(256)
f32_to_bf16_2d_accum_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols,
                            long long ld_in, long long ld_out) {        // out += in   (bf16 += fp32)
  const long long n4 = static_cast<long long>(rows) * (cols / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / (cols / 4)), c = static_cast<int>(i % (cols / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + r * ld_in + c);
    uint2* op = reinterpret_cast<uint2*>(out + r * ld_out + c);
    const uint2 o = *op;
    const float2 o0 = unpack_bf16(o.x), o1 = unpack_bf16(o.y);
    *op = make_uint2(pack_bf16(v.x + o0.x, v.y + o0.y), pack_bf16(v.z + o1.x, v.w + o1.y));
  }
}
}  // namespace qn

// C = A·B with the K loop split `splits` ways over CTAs (for few-tile / long-K shapes such as weight
// gradients).  acc32: fp32 workspace [M, N] that the caller has ZEROED; the bf16 result is written to c.
extern "C" int qn_gemm_bf16_splitk(const void* a, const void* b, void* c, float* acc32, int M, int N, int K,
                                   long long lda, long long ldb, long long ldc, int a_mn, int b_mn, int splits,
                                   int block_n, cudaStream_t stream) {
  using namespace qn;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 4)) return 910;
  if (block_n == 0) block_n = N > 128 ? 256 : 128;
  const int num_kb = (K + 63) / 64;
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  {  // no empty K-slices: every work unit must issue at least one MMA before its accumulator is read
    const int per = (num_kb + splits - 1) / splits;
    splits = (num_kb + per - 1) / per;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.epilogue = EPI_NONE;
  p.splits = splits; p.acc32 = acc32; p.ld32 = N;
  CUtensorMap ta, tb, tcm;
  int rc;
  if (!a_mn) rc = make_tmap(&ta, a, M, K, lda, 64, BLOCK_M); else rc = make_tmap(&ta, a, K, M, lda, 64, BLOCK_K);
  if (rc) return rc;
  const bool two_cta = use_cta_pairs(M, N, block_n, splits, 0);
  if (!b_mn) rc = make_tmap(&tb, b, N, K, ldb, 64, two_cta ? 128 : block_n); else rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K);
  if (rc) return rc;
  if ((rc = make_tmap(&tcm, c, M, N, ldc, 64, 32))) return rc;
#define QN_DISPATCH_SK(BN)                                                                               \
  do {                                                                                                    \
    if (!a_mn && !b_mn) rc = launch<BN, false, false>(ta, tb, tcm, tcm, p, 0, stream);                    \
    else if (!a_mn && b_mn) rc = launch<BN, false, true>(ta, tb, tcm, tcm, p, 0, stream);                 \
    else if (a_mn && !b_mn) rc = launch<BN, true, false>(ta, tb, tcm, tcm, p, 0, stream);                 \
    else rc = launch<BN, true, true>(ta, tb, tcm, tcm, p, 0, stream);                                     \
  } while (0)
  if (two_cta) {
    if (!a_mn && !b_mn) rc = launch_2cta<false, false>(ta, tb, tcm, tcm, p, stream);
    else if (!a_mn && b_mn) rc = launch_2cta<false, true>(ta, tb, tcm, tcm, p, stream);
    else if (a_mn && !b_mn) rc = launch_2cta<true, false>(ta, tb, tcm, tcm, p, stream);
    else rc = launch_2cta<true, true>(ta, tb, tcm, tcm, p, stream);
  } else if (block_n == 256) QN_DISPATCH_SK(256); else QN_DISPATCH_SK(128);
#undef QN_DISPATCH_SK
  if (rc) return rc;
  if (splits > 1) {
    long long n4 = static_cast<long long>(M) * (N / 4);
    int grid = static_cast<int>((n4 + 255) / 256 > 148 * 8 ? 148 * 8 : (n4 + 255) / 256);
    f32_to_bf16_2d_kernel<<<grid, 256, 0, stream>>>(acc32, static_cast<__nv_bfloat16*>(c), M, N, N, ldc);
    return static_cast<int>(cudaGetLastError());
  }
  return 0;
}

// ------------------------------------------------------------------------------- fused TP entry points
namespace qn {
// CTA pairs inside the fused AG / RS kernels: default on (QN_TP_2CTA=0 falls back to single-CTA tiles)
static bool tp_cta_pairs() {
  const char* e = getenv("QN_TP_2CTA");
  return e ? atoi(e) != 0 : true;
}
static int fill_comm(TpComm& c, int rank, int world, unsigned seq, int M) {
  if (world < 2 || world > TP_MAX || M % (128 * world)) return 915;
  c.rank = rank; c.world = world; c.seq = seq;
  c.rows_per_rank = M / world; c.mb_per_rank = c.rows_per_rank / 128;
  const char* dbg = getenv("QN_TP_DEBUG");
  c.debug = dbg ? static_cast<uint32_t>(atoi(dbg)) : 0u;
  const char* mode = getenv("QN_TP_AG_MODE");                // "pump" (default) | "lsu" | "ce" (copy engine, Python side)
  c.ag_mode = (mode && mode[0] == 'c') ? AG_EXTERNAL : (mode && mode[0] == 'l') ? AG_LSU : AG_PUMP;
  // Default: ONE done-flag per source.  Publishing the shard in kAgParts parts (QN_TP_AG_PARTS=4) lets the consumer
  // start on the first rows while the rest is in flight, but every part costs each pump a store-group drain plus a
  // system fence, and the pump is the critical path: measured slower (c_attn 75.6 vs 66.9 us, tp2 step 473 vs 455 ms;
  // profiles/r2_tp_kernel_times.json vs r2_tp_kernel_times_1part.json).
  const char* parts = getenv("QN_TP_AG_PARTS");
  c.ag_parts = (c.mb_per_rank % kAgParts == 0 && parts && atoi(parts) == static_cast<int>(kAgParts)) ? kAgParts : 1u;
  c.rs_parts = 1;                                            // set by the RS launcher once the tile grouping is known
  return 0;
}
// m-blocks per rasterisation group: the largest of 8/4/2/1 that divides the tile rows of one rank, so that no group
// straddles an ownership boundary (remote-before-own ordering of FUSE_RS / own-before-remote of FUSE_AG stays strict
// for every persistent worker — a straddling group could interleave waits and sends of two ranks into a cycle)
static uint32_t fused_group_m(uint32_t tile_rows_per_rank) {
  for (uint32_t g = 8; g > 1; g >>= 1) if (tile_rows_per_rank % g == 0) return g;
  return 1;
}
}  // namespace qn

// K3  all-gather -> GEMM:  C[M, N] = epilogue( AG(x_local)[M, K] · B )
//   x_local [M/world, K] contiguous; gather[r] / ag_flags[r]: rank r's symmetric gather buffer [M, K] and flags [M/32]
//   (host arrays of `world` peer-mapped pointers); b as in qn_gemm_bf16; seq: per-workspace launch counter (>= 1).
//   x_full (optional, pump mode only): private [M, K] tensor that receives the complete gathered operand.
extern "C" int qn_gemm_ag_bf16(const void* x_local, void* const* gather, void* const* ag_flags, int rank, int world,
                               unsigned seq, const void* b, void* c, int M, int N, int K, long long ldb,
                               long long ldc, int b_mn, int epilogue, const void* bias, void* z, long long ldz,
                               int block_n, int max_ctas, void* x_full, cudaStream_t stream) {
  using namespace qn;
  TpComm comm{};
  if (int rc = fill_comm(comm, rank, world, seq, M)) return rc;
  if ((K % 8) || (ldb % 8) || (ldc % 8)) return 910;
  comm.x_local = static_cast<const __nv_bfloat16*>(x_local);
  if (reinterpret_cast<uintptr_t>(x_local) & 15) return 910;
  if (x_full != nullptr && (comm.ag_mode != AG_PUMP || (reinterpret_cast<uintptr_t>(x_full) & 15))) return 918;
  comm.x_full = static_cast<__nv_bfloat16*>(x_full);
  for (int r = 0; r < world; ++r) {
    comm.gather[r] = static_cast<__nv_bfloat16*>(gather[r]);
    comm.ag_flags[r] = static_cast<uint32_t*>(ag_flags[r]);
  }
  if (block_n == 0) block_n = (N > 128 && static_cast<long long>(M / 128) * ((N + 255) / 256) >= 96) ? 256 : 128;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.epilogue = epilogue;
  p.has_z = (epilogue == EPI_BIAS_GELU && z != nullptr) ? 1 : 0;
  if (epilogue == EPI_MUL_AUX && (z == nullptr || (ldz % 8))) return 916;
  CUtensorMap ta, tb, tcm, tz, tx;
  int rc;
  if ((rc = make_tmap(&ta, x_local, comm.rows_per_rank, K, K, 64, BLOCK_M))) return rc;
  if ((rc = make_tmap(&tx, gather[rank], M, K, K, 64, BLOCK_M))) return rc;
  const bool two_cta = tp_cta_pairs() && (comm.rows_per_rank % 256) == 0 && use_cta_pairs(M, N, block_n, 1, 0);
  p.group_m = fused_group_m(comm.mb_per_rank / (two_cta ? 2 : 1));
  if (!b_mn) rc = make_tmap(&tb, b, N, K, ldb, 64, two_cta ? 128 : block_n); else rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K);
  if (rc) return rc;
  if ((rc = make_tmap(&tcm, c, M, N, ldc, 64, 32))) return rc;
  if (p.has_z || epilogue == EPI_MUL_AUX) { if ((rc = make_tmap(&tz, z, M, N, ldz, 64, 32))) return rc; } else tz = tcm;
  if (two_cta) {
    if (!b_mn) return launch_2cta<false, false, FUSE_AG>(ta, tb, tcm, tz, p, stream, &tx, &comm, max_ctas);
    return launch_2cta<false, true, FUSE_AG>(ta, tb, tcm, tz, p, stream, &tx, &comm, max_ctas);
  }
  if (block_n == 256) {
    if (!b_mn) return launch<256, false, false, FUSE_AG>(ta, tb, tcm, tz, p, max_ctas, stream, &tx, &comm);
    return launch<256, false, true, FUSE_AG>(ta, tb, tcm, tz, p, max_ctas, stream, &tx, &comm);
  }
  if (!b_mn) return launch<128, false, false, FUSE_AG>(ta, tb, tcm, tz, p, max_ctas, stream, &tx, &comm);
  return launch<128, false, true, FUSE_AG>(ta, tb, tcm, tz, p, max_ctas, stream, &tx, &comm);
}

// K4  GEMM -> reduce-scatter:  C_local[M/world, N] = sum_ranks( A[M, K] · B ) rows of this rank (+ bias)
//   stage[r] / rs_flags[r]: rank r's symmetric staging buffer — [world][M/world/32 * ceil(N/64)] slab images of 4 KB —
//   and flags [(M/world/128) * 4 * ceil(N/64) * world] (host arrays of peer-mapped pointers).
extern "C" int qn_gemm_rs_bf16(const void* a, const void* b, void* c_local, void* const* stage, void* const* rs_flags,
                               int rank, int world, unsigned seq, int M, int N, int K, long long lda, long long ldb,
                               long long ldc, int b_mn, const void* bias, int block_n, int max_ctas,
                               cudaStream_t stream) {
  using namespace qn;
  TpComm comm{};
  if (int rc = fill_comm(comm, rank, world, seq, M)) return rc;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 8)) return 910;
  for (int r = 0; r < world; ++r) {
    comm.stage[r] = static_cast<__nv_bfloat16*>(stage[r]);
    comm.rs_flags[r] = static_cast<uint32_t*>(rs_flags[r]);
  }
  if (block_n == 0) block_n = (N > 128 && static_cast<long long>(M / 128) * ((N + 255) / 256) >= 96) ? 256 : 128;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.epilogue = bias ? EPI_BIAS : EPI_NONE;
  CUtensorMap ta, tb, tcm;
  int rc;
  if ((rc = make_tmap(&ta, a, M, K, lda, 64, BLOCK_M))) return rc;
  const bool two_cta = tp_cta_pairs() && (comm.rows_per_rank % 256) == 0 && use_cta_pairs(M, N, block_n, 1, 0);
  p.group_m = fused_group_m(comm.mb_per_rank / (two_cta ? 2 : 1));
  {
    // row parts must be whole tile groups, so that the (owner, part) key is monotone along every warp's tile loop
    const uint32_t tr = comm.mb_per_rank / (two_cta ? 2 : 1);
    const char* parts = getenv("QN_TP_RS_PARTS");
    comm.rs_parts = (tr % kAgParts == 0 && (tr / kAgParts) % p.group_m == 0 && parts && atoi(parts) == static_cast<int>(kAgParts))
                        ? kAgParts : 1u;   // opt-in (QN_TP_RS_PARTS=4): measured slower than one flag per owner, see fill_comm
  }
  if (!b_mn) rc = make_tmap(&tb, b, N, K, ldb, 64, two_cta ? 128 : block_n); else rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K);
  if (rc) return rc;
  if ((rc = make_tmap(&tcm, c_local, comm.rows_per_rank, N, ldc, 64, 32))) return rc;
  const CUtensorMap& tx = tcm;                 // unused by FUSE_RS (staged slabs move as raw 4 KB bulk copies)
  if (two_cta) {
    if (!b_mn) return launch_2cta<false, false, FUSE_RS>(ta, tb, tcm, tcm, p, stream, &tx, &comm, max_ctas);
    return launch_2cta<false, true, FUSE_RS>(ta, tb, tcm, tcm, p, stream, &tx, &comm, max_ctas);
  }
  if (block_n == 256) {
    if (!b_mn) return launch<256, false, false, FUSE_RS>(ta, tb, tcm, tcm, p, max_ctas, stream, &tx, &comm);
    return launch<256, false, true, FUSE_RS>(ta, tb, tcm, tcm, p, max_ctas, stream, &tx, &comm);
  }
  if (!b_mn) return launch<128, false, false, FUSE_RS>(ta, tb, tcm, tcm, p, max_ctas, stream, &tx, &comm);
  return launch<128, false, true, FUSE_RS>(ta, tb, tcm, tcm, p, max_ctas, stream, &tx, &comm);
}

// Run-time deadline (in SM cycles) of the peer-flag waits in this translation unit; see common.cuh.
extern "C" int qn_gemm_set_flag_watchdog(long long cycles) {
  return static_cast<int>(cudaMemcpyToSymbol(qn::g_flag_watchdog_cycles, &cycles, sizeof(cycles)));
}

// FP8 forward GEMM with row-wise ("1 x K block") scaling:  C[M, N] (bf16) = epilogue( (A8 · B8ᵀ) ⊙ (sa ⊗ sb) )
//   a8 [M, K] and b8 [N, K] hold e4m3 bytes, K-major, leading dimensions in bytes (multiples of 16);
//   scale_a [M], scale_b [N] fp32 dequantisation scales (qn_quantize_rowwise_e4m3 produces both tensors).
//   epilogue: EPI_NONE / EPI_BIAS / EPI_BIAS_GELU (z = gelu'(pre-activation) as in the bf16 kernel).
extern "C" int qn_gemm_fp8(const void* a8, const void* b8, void* c, int M, int N, int K, long long lda, long long ldb,
                           long long ldc, const float* scale_a, const float* scale_b, int epilogue, const void* bias,
                           void* z, long long ldz, int block_n, cudaStream_t stream) {
  using namespace qn;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % 16) || (lda % 16) || (ldb % 16) || (ldc % 8) || (N % 4) || (reinterpret_cast<uintptr_t>(a8) & 15) ||
      (reinterpret_cast<uintptr_t>(b8) & 15) || (reinterpret_cast<uintptr_t>(c) & 15) ||
      (reinterpret_cast<uintptr_t>(scale_b) & 15))
    return 910;
  if (epilogue != EPI_NONE && epilogue != EPI_BIAS && epilogue != EPI_BIAS_GELU) return 911;
  if (block_n == 0) {
    const long long t256 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256);
    block_n = (N > 128 && t256 >= 96) ? 256 : 128;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K / 2;                       // byte pairs: the kernel's K loop walks 128-byte tile rows
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.epilogue = epilogue;
  p.has_z = (epilogue == EPI_BIAS_GELU && z != nullptr) ? 1 : 0;
  p.fp8 = 1; p.scale_a = scale_a; p.scale_b = scale_b;
  if (epilogue != EPI_NONE && bias != nullptr && (N % 8)) return 912;
  CUtensorMap ta, tb, tc, tz;
  int rc;
  // CTA pairs for fp8: same kernel path as bf16 pairs with kind::f8f6f4.cta_group::2 — opt-in (QN_FP8_2CTA=1) until
  // it has been validated on hardware
  static const bool fp8_pairs = [] { const char* e = getenv("QN_FP8_2CTA"); return e && atoi(e) != 0; }();
  const bool two_cta = fp8_pairs && use_cta_pairs(M, N, block_n, 1, 0);
  if ((rc = make_tmap(&ta, a8, M, K / 2, lda / 2, 64, BLOCK_M))) return rc;
  if ((rc = make_tmap(&tb, b8, N, K / 2, ldb / 2, 64, two_cta ? 128 : block_n))) return rc;
  if ((rc = make_tmap(&tc, c, M, N, ldc, 64, 32))) return rc;
  if (p.has_z) {
    if ((ldz % 8) || (reinterpret_cast<uintptr_t>(z) & 15)) return 913;
    if ((rc = make_tmap(&tz, z, M, N, ldz, 64, 32))) return rc;
  } else {
    tz = tc;
  }
  if (two_cta) return launch_2cta<false, false, FUSE_NONE, false, true>(ta, tb, tc, tz, p, stream);
  if (block_n == 256) return launch<256, false, false, FUSE_NONE, false, true>(ta, tb, tc, tz, p, 0, stream);
  return launch<128, false, false, FUSE_NONE, false, true>(ta, tb, tc, tz, p, 0, stream);
}

// Publish `value` at a (peer-mapped) flag after everything earlier in `stream` — e.g. a cudaMemcpyAsync into the
// peer's gather buffer — has completed: the external-push variant of the all-gather (QN_TP_AG_MODE=ce).
namespace qn {
__global__ void write_flag_kernel(uint32_t* flag, uint32_t value) {
  __threadfence_system();
  st_release_sys(flag, value);
}
}  // namespace qn
extern "C" int qn_write_flag(void* flag_base, long long index, unsigned value, cudaStream_t stream) {
  qn::write_flag_kernel<<<1, 1, 0, stream>>>(static_cast<uint32_t*>(flag_base) + index, value);
  return static_cast<int>(cudaGetLastError());
}
extern "C" int qn_ag_done_base() { return static_cast<int>(qn::kAgDoneBase); }

// Weight-gradient accumulation into an fp32 MAIN-GRAD buffer:  G[M, N] (fp32) += Aᵀ-stored · B-stored  (a: [K, M],
// b: [K, N], both MN-major).  Every work unit (tile x K-slice) adds its TMEM accumulator straight into G with vector
// red.global.add.f32 — the same epilogue split-K uses — so micro-batch gradients accumulate in fp32 with no zeroing,
// no conversion pass and no elementwise add, and split-K comes for free.
extern "C" int qn_gemm_bf16_wgrad_accum_f32(const void* a, const void* b, float* g32, int M, int N, int K, long long lda,
                                            long long ldb, long long ldg, int splits, int block_n, cudaStream_t stream) {
  using namespace qn;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldg % 4) || (N % 4) || (reinterpret_cast<uintptr_t>(g32) & 15)) return 910;
  if (block_n == 0) block_n = N > 128 ? 256 : 128;
  const int num_kb = (K + 63) / 64;
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  {
    const int per = (num_kb + splits - 1) / splits;
    splits = (num_kb + per - 1) / per;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.epilogue = EPI_NONE;
  p.splits = splits; p.acc32 = g32; p.ld32 = static_cast<uint32_t>(ldg); p.acc_f32 = 1;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap(&ta, a, K, M, lda, 64, BLOCK_K))) return rc;
  if ((rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K))) return rc;
  if (use_cta_pairs(M, N, block_n, splits, 0)) return launch_2cta<true, true>(ta, tb, ta, ta, p, stream);
  if (block_n == 256) return launch<256, true, true>(ta, tb, ta, ta, p, 0, stream);
  return launch<128, true, true>(ta, tb, ta, ta, p, 0, stream);
}

// Weight-gradient accumulation:  C[M, N] += Aᵀ-stored · B-stored  (a: [K, M], b: [K, N], both MN-major — the wgrad
// layout dW += dYᵀ · X).  The epilogue reloads the old C tile by TMA, so micro-batch gradients accumulate straight into
// the gradient buffer without a separate elementwise pass.  splits > 1: split-K into the zeroed fp32 workspace
// `acc32`, then one pass adds it into c.
extern "C" int qn_gemm_bf16_wgrad_accum(const void* a, const void* b, void* c, float* acc32, int M, int N, int K,
                                        long long lda, long long ldb, long long ldc, int splits, int block_n,
                                        cudaStream_t stream) {
  using namespace qn;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || (N % 4) || (reinterpret_cast<uintptr_t>(c) & 15)) return 910;
  if (block_n == 0) block_n = N > 128 ? 256 : 128;
  const int num_kb = (K + 63) / 64;
  if (splits > num_kb) splits = num_kb;
  if (splits < 1 || acc32 == nullptr) splits = 1;
  if (splits > 1) {
    const int per = (num_kb + splits - 1) / splits;
    splits = (num_kb + per - 1) / per;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.epilogue = EPI_NONE;
  p.splits = splits; p.acc32 = acc32; p.ld32 = N;
  CUtensorMap ta, tb, tcm;
  int rc;
  if ((rc = make_tmap(&ta, a, K, M, lda, 64, BLOCK_K))) return rc;
  if ((rc = make_tmap(&tb, b, K, N, ldb, 64, BLOCK_K))) return rc;
  if ((rc = make_tmap(&tcm, c, M, N, ldc, 64, 32))) return rc;
  const bool two_cta = use_cta_pairs(M, N, block_n, splits, 0);
  if (splits > 1) {                 // partial sums go to acc32 (the ACCUM epilogue is not used on this path)
    if (two_cta) rc = launch_2cta<true, true>(ta, tb, tcm, tcm, p, stream);
    else if (block_n == 256) rc = launch<256, true, true>(ta, tb, tcm, tcm, p, 0, stream);
    else rc = launch<128, true, true>(ta, tb, tcm, tcm, p, 0, stream);
    if (rc) return rc;
    long long n4 = static_cast<long long>(M) * (N / 4);
    int grid = static_cast<int>((n4 + 255) / 256 > 148 * 8 ? 148 * 8 : (n4 + 255) / 256);
    f32_to_bf16_2d_accum_kernel<<<grid, 256, 0, stream>>>(acc32, static_cast<__nv_bfloat16*>(c), M, N, N, ldc);
    return static_cast<int>(cudaGetLastError());
  }
  if (two_cta) return launch_2cta<true, true, FUSE_NONE, true>(ta, tb, tcm, tcm, p, stream);
  if (block_n == 256) return launch<256, true, true, FUSE_NONE, true>(ta, tb, tcm, tcm, p, 0, stream);
  return launch<128, true, true, FUSE_NONE, true>(ta, tb, tcm, tcm, p, 0, stream);
}
