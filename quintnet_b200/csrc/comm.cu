// quintnet_b200 — in-kernel collectives over NVLink peer memory / NVLS multicast (sm_100a).
//
// Every buffer is "symmetric": the same allocation exists on every rank of the group and each rank has all
// peers' device pointers mapped (ops/symm.py).  The kernels issue the communication themselves:
//   K1  all_reduce      : two-shot — rank r reduces slice r from every peer (fp32 accumulation) and stores
//                         the result straight into every peer's buffer; or, with an NVLS multicast mapping,
//                         one multimem.ld_reduce + multimem.st per 16 bytes (reduction inside the NVSwitch).
//   K2a reduce_scatter  : + 1/P scale + masked squared-norm partial (ZeRO-1 phase A)
//   K2b adamw_allgather : AdamW on the local fp32 shard, updated bf16/fp32 parameters pushed into every
//                         peer's parameter buffer (ZeRO-1 phase B)
// Cross-rank ordering uses per-block release/acquire flags in a symmetric signal pad with a monotonically
// increasing epoch (no host round-trips, no NCCL).  A watchdog turns a lost peer into a trap.
#include "common.cuh"

namespace qn {

constexpr int COMM_MAX_RANKS = 8;
constexpr int COMM_THREADS = 512;

struct PeerPtrs {
  void* buf[COMM_MAX_RANKS];        // peer r's mapping of the data buffer (same offset applied by caller)
  uint32_t* sig[COMM_MAX_RANKS];    // peer r's signal pad: [num_blocks][COMM_MAX_RANKS] uint32
  void* mc;                         // NVLS multicast address of the data buffer (or null)
  int rank, world;
};

// All ranks' block `blockIdx.x` meet here.  slot layout: sig[dst][block][src].
QN_DEVICE void block_barrier(const PeerPtrs& pp, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < pp.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_relaxed_sys(pp.sig[peer] + blockIdx.x * COMM_MAX_RANKS + pp.rank, epoch);   // fence above = release
    wait_flag_ge(pp.sig[pp.rank] + blockIdx.x * COMM_MAX_RANKS + peer, epoch);
  }
  __syncthreads();
}

template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  static QN_DEVICE void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  static QN_DEVICE uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  static QN_DEVICE uint4 mc_reduce(const void* p) { return multimem_ld_reduce_f32x4(p); }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 8;
  static QN_DEVICE void unpack(const uint4& u, float (&f)[8]) {
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static QN_DEVICE uint4 pack(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
  static QN_DEVICE uint4 mc_reduce(const void* p) { return multimem_ld_reduce_bf16x8(p); }
};

// ------------------------------------------------------------------------------------------------ K1
// In-place TWO-SHOT all-reduce(+scale) of n elements (n % (world * VecN) == 0): rank r reduces slice r (NVLS: one
// multimem.ld_reduce per 16 bytes, the sum is formed inside the NVSwitch; P2P: one load per peer, fp32 accumulation)
// and writes the result to every rank (NVLS: one multimem.st; P2P: one store per peer).  kUnroll independent
// vectors per thread are in flight: a peer / switch round trip is ~2-4 us, so per-thread bytes in flight — not the
// thread count — set the achieved bandwidth (16 K threads x 16 B x 1 would cap at ~75 GB/s).
constexpr int kArUnroll = 8;
template <typename T, bool USE_MC>
__global__ void __launch_bounds__(COMM_THREADS)
allreduce_kernel(PeerPtrs pp, long long n, float scale, uint32_t epoch) {
  constexpr int VN = V16<T>::N;
  constexpr int U = USE_MC ? kArUnroll : 4;                  // P2P already has `world` loads in flight per vector
  block_barrier(pp, epoch);                                  // every rank's input is complete
  const long long nvec = n / VN;
  const long long per_rank = nvec / pp.world;
  const long long lo = per_rank * pp.rank;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < per_rank; i0 += stride * U) {
    uint4 red[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < per_rank) {
        const long long v = lo + i;
        if constexpr (USE_MC) {
          red[u] = V16<T>::mc_reduce(reinterpret_cast<const uint4*>(pp.mc) + v);
        } else {
          float acc[VN];
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] = 0.f;
#pragma unroll
          for (int p = 0; p < COMM_MAX_RANKS; ++p) {
            if (p < pp.world) {
              const int src = (pp.rank + p) % pp.world;      // stagger peers across ranks
              float f[VN];
              V16<T>::unpack(ld_cv_v4(reinterpret_cast<const uint4*>(pp.buf[src]) + v), f);
#pragma unroll
              for (int e = 0; e < VN; ++e) acc[e] += f[e];
            }
          }
          red[u] = V16<T>::pack(acc);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < per_rank) {
        const long long v = lo + i;
        uint4 out = red[u];
        if (scale != 1.0f) {
          float acc[VN];
          V16<T>::unpack(out, acc);
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] *= scale;
          out = V16<T>::pack(acc);
        }
        if constexpr (USE_MC) {
          multimem_st_v4(reinterpret_cast<uint4*>(pp.mc) + v, out);
        } else {
#pragma unroll
          for (int p = 0; p < COMM_MAX_RANKS; ++p)
            if (p < pp.world) st_v4(reinterpret_cast<uint4*>(pp.buf[(pp.rank + p) % pp.world]) + v, out);
        }
      }
    }
  }
  block_barrier(pp, epoch + 1);                              // every rank's writes have landed everywhere
}

// ONE-SHOT out-of-place all-reduce for small messages (TP activations of the latency-bound ViT path, scalars):
//   1. every rank copies its input into its half of a double-buffered symmetric staging area,
//   2. ONE cross-rank barrier (per block),
//   3. every rank reduces ALL n elements itself (NVLS: multimem.ld_reduce; P2P: a load per peer) into its private output.
// One NVLink round trip instead of two, no second barrier (the parity of the per-block launch counter selects the
// staging half, so a fast rank's next launch cannot overwrite data a slow rank is still reading).  The launch
// counter lives in DEVICE memory and is advanced by the kernel itself, so the launch is CUDA-graph-replayable.
template <typename T, bool USE_MC>
__global__ void __launch_bounds__(COMM_THREADS)
allreduce_oneshot_kernel(PeerPtrs pp, const T* __restrict__ in, T* __restrict__ out, long long n, float scale,
                         uint32_t* __restrict__ counters, long long half_bytes) {
  constexpr int VN = V16<T>::N;
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = ++counters[blockIdx.x];    // same launch sequence on every rank -> same value
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const long long half = (epoch & 1u) * half_bytes;
  const long long nvec = n / VN;                             // n % VN == 0 (the wrapper pads)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  uint4* mine = reinterpret_cast<uint4*>(static_cast<char*>(pp.buf[pp.rank]) + half);
  for (long long v = first; v < nvec; v += stride) mine[v] = reinterpret_cast<const uint4*>(in)[v];
  // one-shot flags live at uint32 index 1024.. of the signal pad (byte 4096; the two-shot kernels use [0, 64*8) and
  // torch's default pad is 9216 bytes)
  __syncthreads();
  if (threadIdx.x < pp.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_relaxed_sys(pp.sig[peer] + 1024 + blockIdx.x * COMM_MAX_RANKS + pp.rank, epoch);
    wait_flag_ge(pp.sig[pp.rank] + 1024 + blockIdx.x * COMM_MAX_RANKS + peer, epoch);
  }
  __syncthreads();
  for (long long v = first; v < nvec; v += stride) {
    float acc[VN];
    if constexpr (USE_MC) {
      V16<T>::unpack(V16<T>::mc_reduce(reinterpret_cast<const uint4*>(static_cast<const char*>(pp.mc) + half) + v), acc);
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] = 0.f;
#pragma unroll
      for (int p = 0; p < COMM_MAX_RANKS; ++p) {
        if (p < pp.world) {
          float f[VN];
          V16<T>::unpack(ld_cv_v4(reinterpret_cast<const uint4*>(static_cast<const char*>(pp.buf[(pp.rank + p) % pp.world]) + half) + v), f);
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] += f[e];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] *= scale;
    reinterpret_cast<uint4*>(out)[v] = V16<T>::pack(acc);
  }
}

// ----------------------------------------------------------------------------------------------- K2a
// out[i] = scale * sum_p grad_p[lo + i]  (fp32), i in this rank's shard;  sq += sum (out*mask)^2
// USE_MC: the sum is one multimem.ld_reduce per 16 bytes (formed in the NVSwitch); otherwise one load per peer.
template <typename T, bool USE_MC>
__global__ void __launch_bounds__(COMM_THREADS)
reduce_scatter_sqnorm_kernel(PeerPtrs pp, float* __restrict__ out, const float* __restrict__ mask,
                             float* __restrict__ sq, long long n, float scale, uint32_t epoch) {
  constexpr int VN = V16<T>::N;
  constexpr int U = USE_MC ? kArUnroll : 4;
  __shared__ float red[COMM_THREADS / 32];
  block_barrier(pp, epoch);
  const long long per_rank = n / pp.world;                   // elements
  const long long nvec = per_rank / VN;
  const long long lo_vec = (per_rank * pp.rank) / VN;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float local = 0.f;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
    uint4 raw[U];
    float accs[U][VN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        if constexpr (USE_MC) {
          raw[u] = V16<T>::mc_reduce(reinterpret_cast<const uint4*>(pp.mc) + lo_vec + i);
        } else {
#pragma unroll
          for (int e = 0; e < VN; ++e) accs[u][e] = 0.f;
#pragma unroll
          for (int p = 0; p < COMM_MAX_RANKS; ++p) {
            if (p < pp.world) {
              const int src = (pp.rank + p) % pp.world;
              float f[VN];
              V16<T>::unpack(ld_cv_v4(reinterpret_cast<const uint4*>(pp.buf[src]) + lo_vec + i), f);
#pragma unroll
              for (int e = 0; e < VN; ++e) accs[u][e] += f[e];
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        if constexpr (USE_MC) V16<T>::unpack(raw[u], accs[u]);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float a = accs[u][e] * scale;
          const float mk = mask != nullptr ? mask[i * VN + e] : 1.0f;
          local += a * a * mk;
          out[i * VN + e] = a;
        }
      }
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < COMM_THREADS / 32; ++w) t += red[w];
    atomicAdd(sq, t);
  }
  block_barrier(pp, epoch + 1);                              // peers are done reading my gradient buffer
}

// ----------------------------------------------------------------------------------------------- K2b
// AdamW on the local shard (fp32 master/m/v, fp32 reduced gradient) and all-gather of the new parameters:
// the updated value is stored into every peer's parameter buffer at [lo + i].
// USE_MC: the updated parameters go to every rank with ONE multimem.st per 16 bytes (the NVSwitch replicates).
template <typename PT, bool USE_MC>
__global__ void __launch_bounds__(COMM_THREADS)
adamw_allgather_kernel(PeerPtrs pp, const float* __restrict__ grad, float* __restrict__ master, float* __restrict__ m,
                       float* __restrict__ v, const float* __restrict__ grad_scale, long long n, float lr, float beta1,
                       float beta2, float eps, float wd, float inv_bc1, float inv_sqrt_bc2, uint32_t epoch) {
  const float gs = grad_scale != nullptr ? *grad_scale : 1.0f;
  const long long per_rank = n / pp.world;
  const long long lo = per_rank * pp.rank;
  constexpr int EPT = sizeof(PT) == 2 ? 8 : 4;                // elements per thread-iteration = one 16-byte parameter vector
  const long long nv = per_rank / EPT;                        // (per_rank % 8 == 0: checked by the launcher)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float wv[EPT];
#pragma unroll
    for (int h = 0; h < EPT / 4; ++h) {
      const long long o4 = i * EPT + h * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(grad + o4);
      float4 w4 = *reinterpret_cast<const float4*>(master + o4);
      float4 m4 = *reinterpret_cast<const float4*>(m + o4);
      float4 v4 = *reinterpret_cast<const float4*>(v + o4);
      const float* gp = reinterpret_cast<const float*>(&g4);
      float* wp = reinterpret_cast<float*>(&w4);
      float* mp = reinterpret_cast<float*>(&m4);
      float* vp = reinterpret_cast<float*>(&v4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gr = gp[e] * gs;
        mp[e] = beta1 * mp[e] + (1.0f - beta1) * gr;
        vp[e] = beta2 * vp[e] + (1.0f - beta2) * gr * gr;
        wp[e] = wp[e] * (1.0f - lr * wd) - lr * (mp[e] * inv_bc1) / (sqrtf(vp[e]) * inv_sqrt_bc2 + eps);
        wv[h * 4 + e] = wp[e];
      }
      *reinterpret_cast<float4*>(master + o4) = w4;
      *reinterpret_cast<float4*>(m + o4) = m4;
      *reinterpret_cast<float4*>(v + o4) = v4;
    }
    uint4 o;
    if constexpr (sizeof(PT) == 2) {
      o = make_uint4(pack_bf16(wv[0], wv[1]), pack_bf16(wv[2], wv[3]), pack_bf16(wv[4], wv[5]), pack_bf16(wv[6], wv[7]));
    } else {
      o = make_uint4(__float_as_uint(wv[0]), __float_as_uint(wv[1]), __float_as_uint(wv[2]), __float_as_uint(wv[3]));
    }
    const long long byte_off = (lo + i * EPT) * static_cast<long long>(sizeof(PT));
    if constexpr (USE_MC) {
      multimem_st_v4(static_cast<char*>(pp.mc) + byte_off, o);
    } else {
#pragma unroll
      for (int p = 0; p < COMM_MAX_RANKS; ++p)
        if (p < pp.world) st_v4(static_cast<char*>(pp.buf[(pp.rank + p) % pp.world]) + byte_off, o);
    }
  }
  block_barrier(pp, epoch);                                  // all shards have been pushed to all ranks
}

static int fill(PeerPtrs& pp, void* const* bufs, void* const* sigs, void* mc, int rank, int world) {
  if (world < 1 || world > COMM_MAX_RANKS) return 960;
  for (int i = 0; i < world; ++i) { pp.buf[i] = bufs[i]; pp.sig[i] = static_cast<uint32_t*>(sigs[i]); }
  pp.mc = mc; pp.rank = rank; pp.world = world;
  return 0;
}

}  // namespace qn

using namespace qn;

// bufs / sigs: host arrays of `world` device pointers (peer mappings).  The signal pad must hold at least
// blocks * 8 uint32 and `epoch` must grow by >= 2 between launches that use the same pad.
extern "C" int qn_allreduce(void* const* bufs, void* const* sigs, void* mc, int rank, int world, long long n,
                            int dtype, float scale, unsigned int epoch, int blocks, cudaStream_t s) {
  PeerPtrs pp;
  if (int rc = fill(pp, bufs, sigs, mc, rank, world)) return rc;
  const int vn = dtype == 1 ? 8 : 4;
  if (n % (world * vn)) return 961;
  if (mc != nullptr) {
    if (dtype == 1) allreduce_kernel<__nv_bfloat16, true><<<blocks, COMM_THREADS, 0, s>>>(pp, n, scale, epoch);
    else allreduce_kernel<float, true><<<blocks, COMM_THREADS, 0, s>>>(pp, n, scale, epoch);
  } else {
    if (dtype == 1) allreduce_kernel<__nv_bfloat16, false><<<blocks, COMM_THREADS, 0, s>>>(pp, n, scale, epoch);
    else allreduce_kernel<float, false><<<blocks, COMM_THREADS, 0, s>>>(pp, n, scale, epoch);
  }
  return (int)cudaGetLastError();
}

extern "C" int qn_reduce_scatter_sqnorm(void* const* bufs, void* const* sigs, void* mc, int rank, int world, float* out,
                                        const float* mask, float* sq, long long n, int dtype, float scale,
                                        unsigned int epoch, int blocks, cudaStream_t s) {
  PeerPtrs pp;
  if (int rc = fill(pp, bufs, sigs, mc, rank, world)) return rc;
  const int vn = dtype == 1 ? 8 : 4;
  if (n % (world * vn)) return 961;
  if (mc != nullptr) {
    if (dtype == 1) reduce_scatter_sqnorm_kernel<__nv_bfloat16, true><<<blocks, COMM_THREADS, 0, s>>>(pp, out, mask, sq, n, scale, epoch);
    else reduce_scatter_sqnorm_kernel<float, true><<<blocks, COMM_THREADS, 0, s>>>(pp, out, mask, sq, n, scale, epoch);
  } else {
    if (dtype == 1) reduce_scatter_sqnorm_kernel<__nv_bfloat16, false><<<blocks, COMM_THREADS, 0, s>>>(pp, out, mask, sq, n, scale, epoch);
    else reduce_scatter_sqnorm_kernel<float, false><<<blocks, COMM_THREADS, 0, s>>>(pp, out, mask, sq, n, scale, epoch);
  }
  return (int)cudaGetLastError();
}

// One-shot out-of-place all-reduce of a SMALL message (n * sizeof(T) <= half_bytes): `stage`/`sigs`/`mc` describe the
// symmetric staging buffer (2 * half_bytes), `in`/`out` are private tensors (may alias each other), `counters` is a
// private zero-initialised uint32[blocks] device array that the kernel advances itself (CUDA-graph safe).
extern "C" int qn_allreduce_oneshot(void* const* stage, void* const* sigs, void* mc, int rank, int world, const void* in,
                                    void* out, long long n, int dtype, float scale, unsigned int* counters,
                                    long long half_bytes, int blocks, cudaStream_t s) {
  PeerPtrs pp;
  if (int rc = fill(pp, stage, sigs, mc, rank, world)) return rc;
  const int vn = dtype == 1 ? 8 : 4;
  if (n % vn || n * (dtype == 1 ? 2 : 4) > half_bytes || blocks < 1 || blocks > 64) return 962;
  if (mc != nullptr) {
    if (dtype == 1) allreduce_oneshot_kernel<__nv_bfloat16, true><<<blocks, COMM_THREADS, 0, s>>>(pp, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, scale, counters, half_bytes);
    else allreduce_oneshot_kernel<float, true><<<blocks, COMM_THREADS, 0, s>>>(pp, (const float*)in, (float*)out, n, scale, counters, half_bytes);
  } else {
    if (dtype == 1) allreduce_oneshot_kernel<__nv_bfloat16, false><<<blocks, COMM_THREADS, 0, s>>>(pp, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n, scale, counters, half_bytes);
    else allreduce_oneshot_kernel<float, false><<<blocks, COMM_THREADS, 0, s>>>(pp, (const float*)in, (float*)out, n, scale, counters, half_bytes);
  }
  return (int)cudaGetLastError();
}

extern "C" int qn_adamw_allgather(void* const* bufs, void* const* sigs, void* mc, int rank, int world, const float* grad,
                                  float* master, float* m, float* v, const float* grad_scale, long long n,
                                  int pdtype, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                  float bc2, unsigned int epoch, int blocks, cudaStream_t s) {
  PeerPtrs pp;
  if (int rc = fill(pp, bufs, sigs, mc, rank, world)) return rc;
  if (n % (world * 8)) return 961;
  const float inv_bc1 = 1.0f / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
#define QN_AG_LAUNCH(PT, MC) adamw_allgather_kernel<PT, MC><<<blocks, COMM_THREADS, 0, s>>>(pp, grad, master, m, v, grad_scale, n, lr, \
                                                         beta1, beta2, eps, wd, inv_bc1, inv_sqrt_bc2, epoch)
  if (mc != nullptr) { if (pdtype == 1) QN_AG_LAUNCH(__nv_bfloat16, true); else QN_AG_LAUNCH(float, true); }
  else { if (pdtype == 1) QN_AG_LAUNCH(__nv_bfloat16, false); else QN_AG_LAUNCH(float, false); }
#undef QN_AG_LAUNCH
  return (int)cudaGetLastError();
}

// ---- plain IPC helpers (fallback rendezvous when torch symmetric memory is unavailable)
extern "C" int qn_ipc_get_handle(void* ptr, void* handle64) {
  return (int)cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr);
}
extern "C" int qn_ipc_open_handle(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int qn_malloc(void** out, long long bytes) { return (int)cudaMalloc(out, (size_t)bytes); }
extern "C" int qn_memset0(void* p, long long bytes, cudaStream_t s) { return (int)cudaMemsetAsync(p, 0, (size_t)bytes, s); }

// Run-time deadline (in SM cycles) of the peer-flag waits in this translation unit; see common.cuh.
extern "C" int qn_comm_set_flag_watchdog(long long cycles) {
  return static_cast<int>(cudaMemcpyToSymbol(qn::g_flag_watchdog_cycles, &cycles, sizeof(cycles)));
}

// ---------------------------------------------------------------------------------------------------------
// NVLink push micro-benchmark (scripts/bench_nvlink_push.py): how fast can `blocks x threads` threads copy a
// local buffer into peer memory with 16-byte stores, `DEPTH` loads in flight per thread.  seg_vecs > 0 writes
// `seg_vecs` contiguous vectors and then jumps `seg_stride` vectors (the row-segment pattern of a GEMM epilogue
// scattering a [rows x 64] slab into a wider matrix).
// ---------------------------------------------------------------------------------------------------------
namespace qn {
template <int DEPTH>
__global__ void push_bench_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long nvec, int seg_vecs,
                                  long long seg_stride) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long base = tid; base < nvec; base += nthr * DEPTH) {
    uint4 v[DEPTH];
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      const long long i = base + j * nthr;
      if (i < nvec) v[j] = ld_nc_v4(src + i);
    }
#pragma unroll
    for (int j = 0; j < DEPTH; ++j) {
      const long long i = base + j * nthr;
      if (i < nvec) {
        const long long o = seg_vecs > 0 ? (i / seg_vecs) * seg_stride + (i % seg_vecs) : i;
        st_v4(dst + o, v[j]);
      }
    }
  }
  __threadfence_system();
}
}  // namespace qn

namespace qn {
// Same copy, but staged through shared memory with bulk async copies (one thread per CTA drives it): chunks of
// `chunk` bytes, STAGES deep, loads issued two chunks ahead of the stores.
template <int STAGES>
__global__ void push_bulk_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long bytes, int chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  constexpr int LA = 2;
  const long long nchunks = bytes / chunk;
  const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (long long it = 0; it < mine + LA; ++it) {
    if (it < mine) {
      const int s = it % STAGES;
      tma_store_wait_read<1>();                      // the store that last read this stage has drained it
      mbar_arrive_expect_tx(&full[s], chunk);
      bulk_load_1d(sm + s * chunk, src + (blockIdx.x + it * gridDim.x) * chunk, chunk, &full[s]);
    }
    if (it >= LA) {
      const long long j = it - LA;
      const int s = j % STAGES;
      mbar_wait(&full[s], (j / STAGES) & 1);
      bulk_store_1d(dst + (blockIdx.x + j * gridDim.x) * chunk, sm + s * chunk, chunk);
      tma_store_commit();
    }
  }
  tma_store_wait_all<0>();
  __threadfence_system();
}
}  // namespace qn

extern "C" int qn_push_bench_bulk(const void* src, void* dst, long long bytes, int blocks, int chunk, cudaStream_t s) {
  constexpr int STAGES = 4;
  auto k = qn::push_bulk_kernel<STAGES>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * chunk);
  k<<<blocks, 32, STAGES * chunk, s>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), bytes, chunk);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int qn_push_bench(const void* src, void* dst, long long bytes, int blocks, int threads, int depth,
                             int seg_vecs, long long seg_stride, cudaStream_t s) {
  const long long nvec = bytes / 16;
  const uint4* a = static_cast<const uint4*>(src);
  uint4* b = static_cast<uint4*>(dst);
  if (depth >= 16) qn::push_bench_kernel<16><<<blocks, threads, 0, s>>>(a, b, nvec, seg_vecs, seg_stride);
  else if (depth >= 8) qn::push_bench_kernel<8><<<blocks, threads, 0, s>>>(a, b, nvec, seg_vecs, seg_stride);
  else if (depth >= 4) qn::push_bench_kernel<4><<<blocks, threads, 0, s>>>(a, b, nvec, seg_vecs, seg_stride);
  else qn::push_bench_kernel<1><<<blocks, threads, 0, s>>>(a, b, nvec, seg_vecs, seg_stride);
  return static_cast<int>(cudaGetLastError());
}
