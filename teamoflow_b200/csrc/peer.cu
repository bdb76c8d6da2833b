// Multi-GPU pieces that run directly over NVLink peer memory (one process per GPU, buffers mapped
// into every process with CUDA IPC).  The reference has no distribution at all; these
// kernels are the new-build exchange steps of the two sharded paths:
//
//   * item-sharded top-k:  every rank leaves its per-slab lists [n_users, k] in its own peer buffer;
//     ONE kernel per rank pulls the G lists of the rank's user slice from the peers (P2P loads),
//     merges them with the (score desc, item id asc) comparator and pushes the merged rows into the
//     result buffer of every peer (P2P stores): all-to-all + merge + all-gather in one launch.
//   * user-sharded training:  every rank leaves its partial item-side gradient in its peer buffer;
//     ONE kernel per rank sums its slice of rows over the peers in rank order (deterministic, every
//     replica receives the same bits), optionally applies the Adam step-1 update to that slice of
//     the weights, and pushes the result to every peer: reduce-scatter + update + all-gather.
//
// Ordering between ranks is a stream-ordered flag barrier in peer memory (release/acquire at system
// scope), so no host synchronisation is involved.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace tmf {

constexpr int kMaxPeers = 16;
struct PeerPtrs { void* p[kMaxPeers]; };

__device__ __forceinline__ uint32_t f2key_desc(float x) {  // monotone float -> uint (larger float = larger key)
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f_desc(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------ flag barrier
// pads.p[g] = rank g's pad: kMaxPeers uint32 slots, slot s is written by rank s only.  Epochs only grow.
// A peer that never arrives (crashed rank) trips the timeout: the waiting rank sets the error word of its OWN pad
// (kErrWord, sticky, read by the host at its next synchronisation point -- dist.PeerArena.check()) and leaves the barrier
// instead of hanging the GPU or killing the CUDA context with a trap (a benign host stall on one rank -- GC, a blocked
// print, graph instantiation -- must not take the other ranks' contexts down).
// epoch == 0: the epoch is the next value of a counter kept in the rank's own pad (word kEpochWord, touched by this rank
// only).  Every rank issues the same sequence of barriers, so the counters agree -- and the launch carries no
// per-call argument, which makes it replayable from a CUDA graph.
constexpr int kEpochWord = 32;  // byte 128 of the 256-byte pad; the arrival slots use words [0, kMaxPeers)
constexpr int kErrWord = 33;    // byte 132: set to 1 + (rank that was missing) when a wait timed out
__global__ void peer_barrier_kernel(PeerPtrs pads, int world, int rank, uint32_t epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (epoch == 0) {
    uint32_t e = 0;
    if (t == 0) {
      uint32_t* ctr = reinterpret_cast<uint32_t*>(pads.p[rank]) + kEpochWord;
      e = *ctr + 1u;
      if (e == 0) e = 1u;
      *ctr = e;
    }
    epoch = __shfl_sync(0xffffffffu, e, 0);
  }
  if (t >= world) return;
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(pads.p[t]) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[rank]) + t;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      atomicMax(reinterpret_cast<uint32_t*>(pads.p[rank]) + kErrWord, (uint32_t)(1 + t));
      break;
    }
    __nanosleep(200);
  }
}

// ------------------------------------------------------------------ G-way merge of sorted top-k lists
// One warp per user row.  Every input list is sorted by (score desc, id asc) -- the order tmf_score_topk
// writes -- so the merged rank of entry q of list g is q + sum over the other lists of the number of
// entries that precede it, each found by binary search over 64-bit (score key, ~id) keys in shared memory:
// O(G k log k) per row where counting all pairs costs O((G k)^2).
constexpr int kMergeWarps = 4;

__global__ void __launch_bounds__(kMergeWarps * 32)
topk_merge_lists_kernel(const __grid_constant__ PeerPtrs idx_in, const __grid_constant__ PeerPtrs sc_in, int G, long long row_lo,
                        long long n_rows, int k, const __grid_constant__ PeerPtrs out_idx, const __grid_constant__ PeerPtrs out_sc,
                        int n_out) {  // __grid_constant__: the pointer tables are indexed in parameter space, not copied to the stack
  extern __shared__ __align__(16) unsigned char smraw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long lr = (long long)blockIdx.x * kMergeWarps + w;
  if (lr >= n_rows) return;
  const long long row = row_lo + lr;
  const int n = G * k;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smraw) + (size_t)w * (n + k);
  unsigned long long* outk = keys + n;
  auto pack = [](float sc, int id) {
    return ((unsigned long long)f2key_desc(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)id);
  };
  const bool vec = (k & 3) == 0;  // rows of k entries start 16-byte aligned: 128-bit loads / stores over NVLink
  if (vec) {
    // the loads of ALL lists are issued before the first is consumed: one NVLink round trip per row instead of one per list
    const int k4 = k >> 2;
    for (int q0 = 0; q0 < k4; q0 += 32) {
      const int q4 = q0 + lane;
      constexpr int GB = 8;  // lists per batch (registers: 8 x (int4 + float4))
      for (int g0 = 0; g0 < G; g0 += GB) {
        int4 iv[GB];
        float4 sv[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) {
          if (g0 + j < G && q4 < k4) {  // peer data written by another GPU moments ago: never from a stale line
            iv[j] = __ldcv(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(idx_in.p[g0 + j]) + row * k) + q4);
            sv[j] = __ldcv(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sc_in.p[g0 + j]) + row * k) + q4);
          }
        }
#pragma unroll
        for (int j = 0; j < GB; ++j) {
          if (g0 + j < G && q4 < k4) {
            unsigned long long* dst = keys + (g0 + j) * k + 4 * q4;
            dst[0] = pack(sv[j].x, iv[j].x); dst[1] = pack(sv[j].y, iv[j].y);
            dst[2] = pack(sv[j].z, iv[j].z); dst[3] = pack(sv[j].w, iv[j].w);
          }
        }
      }
    }
  } else {
    for (int g = 0; g < G; ++g) {
      const int* gi = reinterpret_cast<const int*>(idx_in.p[g]) + row * k;
      const float* gs = reinterpret_cast<const float*>(sc_in.p[g]) + row * k;
      for (int q = lane; q < k; q += 32) keys[g * k + q] = pack(__ldcv(gs + q), __ldcv(gi + q));
    }
  }
  __syncwarp();
  for (int t = lane; t < n; t += 32) {
    const int g = t / k, q = t - g * k;
    const unsigned long long key = keys[t];
    int rank = q;
    for (int o = 0; o < G; ++o) {
      if (o == g) continue;
      const unsigned long long* lst = keys + o * k;
      int lo = 0, hi = k;  // first position whose key is not greater than `key`
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] > key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
      if (rank >= k) break;
    }
    if (rank < k) outk[rank] = key;
  }
  __syncwarp();
  auto key_id = [](unsigned long long key) { return (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)); };
  auto key_sc = [](unsigned long long key) { return key2f_desc((uint32_t)(key >> 32)); };
  if (vec) {
    const int k4 = k >> 2;
    for (int q4 = lane; q4 < k4; q4 += 32) {
      const unsigned long long k0 = outk[4 * q4], k1 = outk[4 * q4 + 1], k2 = outk[4 * q4 + 2], k3 = outk[4 * q4 + 3];
      const int4 iv = make_int4(key_id(k0), key_id(k1), key_id(k2), key_id(k3));
      const float4 sv = make_float4(key_sc(k0), key_sc(k1), key_sc(k2), key_sc(k3));
      for (int d = 0; d < n_out; ++d) {
        reinterpret_cast<int4*>(reinterpret_cast<int*>(out_idx.p[d]) + row * k)[q4] = iv;
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out_sc.p[d]) + row * k)[q4] = sv;
      }
    }
  } else {
    for (int d = 0; d < n_out; ++d) {
      int* oi = reinterpret_cast<int*>(out_idx.p[d]) + row * k;
      float* os = reinterpret_cast<float*>(out_sc.p[d]) + row * k;
      for (int q = lane; q < k; q += 32) {
        oi[q] = key_id(outk[q]);
        os[q] = key_sc(outk[q]);
      }
    }
  }
}

// ------------------------------------------------------------------ slice-wise sum over peers (+ Adam step 1) + push
// Rows [row_lo, row_lo + n_rows) of a [*, ld] fp32 matrix: sum of the G peers' partials in rank order, then
//   lr <  0 : the sum is stored into dst of every peer                       (all-reduce)
//   lr >= 0 : w <- adam1(w, sum) for the slice, new rows stored to every peer (all-reduce fused with the update)
// One float4 per thread per trip; the loads of the G partials are issued together.
__device__ __forceinline__ float adam1_update(float w, float g, float lr) { return adam1_apply(w, g, adam1_alpha(lr)); }

template <int G>
__global__ void __launch_bounds__(256) peer_reduce_push_kernel(PeerPtrs part, PeerPtrs dst, long long off4, long long n4,
                                                               int world, int rank, float lr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g < world) v[g] = __ldcv(reinterpret_cast<const float4*>(part.p[g]) + off4 + i);
    }
    float4 s = v[0];
#pragma unroll
    for (int g = 1; g < G; ++g) {
      if (g < world) { s.x += v[g].x; s.y += v[g].y; s.z += v[g].z; s.w += v[g].w; }
    }
    if (lr >= 0.f) {
      const float4 w = *(reinterpret_cast<const float4*>(dst.p[rank]) + off4 + i);
      s.x = adam1_update(w.x, s.x, lr); s.y = adam1_update(w.y, s.y, lr);
      s.z = adam1_update(w.z, s.z, lr); s.w = adam1_update(w.w, s.w, lr);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (g < world) *(reinterpret_cast<float4*>(dst.p[g]) + off4 + i) = s;
    }
  }
}

static int fill_ptrs(PeerPtrs& pp, const void* const* host, int n) {
  for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < n ? const_cast<void*>(host[i]) : nullptr;
  return 0;
}

}  // namespace tmf

using namespace tmf;

// ---- peer buffers: plain cudaMalloc allocations (exportable with the legacy CUDA IPC API), zero-filled
extern "C" int tmf_peer_alloc(size_t bytes, void** out) {
  TMF_REQUIRE(out != nullptr && bytes > 0, "tmf_peer_alloc: bad arguments");
  TMF_CUDA(cudaMalloc(out, bytes));
  TMF_CUDA(cudaMemset(*out, 0, bytes));
  TMF_CUDA(cudaDeviceSynchronize());
  return TMF_OK;
}

extern "C" int tmf_peer_free(void* p) {
  if (p) TMF_CUDA(cudaFree(p));
  return TMF_OK;
}

extern "C" int tmf_ipc_export(const void* dev_ptr, void* handle_host) {
  TMF_REQUIRE(dev_ptr != nullptr && handle_host != nullptr, "tmf_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == TMF_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  TMF_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle_host, &h, sizeof(h));
  return TMF_OK;
}

extern "C" int tmf_ipc_open(const void* handle_host, void** out) {
  TMF_REQUIRE(handle_host != nullptr && out != nullptr, "tmf_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  TMF_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return TMF_OK;
}

extern "C" int tmf_ipc_close(void* p) {
  if (p) TMF_CUDA(cudaIpcCloseMemHandle(p));
  return TMF_OK;
}

extern "C" int tmf_peer_barrier(const void* const* pads_host, int32_t world, int32_t rank, uint32_t epoch, uint32_t timeout_ms,
                                tmf_stream_t stream) {
  TMF_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "tmf_peer_barrier: bad world/rank");
  PeerPtrs pads;
  fill_ptrs(pads, pads_host, world);
  const unsigned long long ms = timeout_ms ? timeout_ms : 120000u;  // 0 = default (2 minutes)
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(pads, world, rank, epoch, ms * 1000ull * 1000ull);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}

static int launch_merge(const PeerPtrs& ii, const PeerPtrs& ss, int G, int64_t row_lo, int64_t n_rows, int k, const PeerPtrs& oi,
                        const PeerPtrs& os, int n_out, cudaStream_t st) {
  if (n_rows <= 0) return TMF_OK;
  const size_t smem = (size_t)kMergeWarps * ((size_t)G * k + k) * 8;
  TMF_REQUIRE(smem <= 200 * 1024, "top-k merge: n_lists*k too large");
  if (smem > 48 * 1024)
    TMF_CUDA(cudaFuncSetAttribute(topk_merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_lists_kernel<<<(unsigned)cdiv(n_rows, kMergeWarps), kMergeWarps * 32, smem, st>>>(ii, ss, G, row_lo, n_rows, k, oi, os,
                                                                                             n_out);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}

extern "C" int tmf_topk_merge(const int32_t* idx_in, const float* score_in, int32_t n_lists, int64_t n_users, int32_t k,
                              int32_t* out_idx, float* out_score, tmf_stream_t stream) {
  TMF_REQUIRE(n_lists >= 1 && n_lists <= kMaxPeers && k >= 1, "tmf_topk_merge: need 1 <= n_lists <= %d, k >= 1", kMaxPeers);
  PeerPtrs ii{}, ss{}, oi{}, os{};
  for (int g = 0; g < n_lists; ++g) {
    ii.p[g] = const_cast<int32_t*>(idx_in) + (size_t)g * n_users * k;
    ss.p[g] = const_cast<float*>(score_in) + (size_t)g * n_users * k;
  }
  oi.p[0] = out_idx;
  os.p[0] = out_score;
  return launch_merge(ii, ss, n_lists, 0, n_users, k, oi, os, 1, as_stream(stream));
}

extern "C" int tmf_topk_merge_peer(const void* const* idx_host, const void* const* score_host, int32_t world, int64_t row_lo,
                                   int64_t n_rows, int32_t k, const void* const* out_idx_host, const void* const* out_score_host,
                                   int32_t n_out, tmf_stream_t stream) {
  TMF_REQUIRE(world >= 1 && world <= kMaxPeers && n_out >= 1 && n_out <= kMaxPeers && k >= 1, "tmf_topk_merge_peer: bad sizes");
  PeerPtrs ii, ss, oi, os;
  fill_ptrs(ii, idx_host, world);
  fill_ptrs(ss, score_host, world);
  fill_ptrs(oi, out_idx_host, n_out);
  fill_ptrs(os, out_score_host, n_out);
  return launch_merge(ii, ss, world, row_lo, n_rows, k, oi, os, n_out, as_stream(stream));
}

extern "C" int tmf_peer_reduce_push(const void* const* part_host, const void* const* dst_host, int32_t world, int32_t rank,
                                    int64_t elem_off, int64_t n_elems, float lr, tmf_stream_t stream) {
  TMF_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "tmf_peer_reduce_push: bad world/rank");
  TMF_REQUIRE(elem_off % 4 == 0 && n_elems % 4 == 0, "tmf_peer_reduce_push: offsets and sizes must be multiples of 4 floats");
  if (n_elems == 0) return TMF_OK;
  PeerPtrs part, dst;
  fill_ptrs(part, part_host, world);
  fill_ptrs(dst, dst_host, world);
  for (int g = 0; g < world; ++g)
    TMF_REQUIRE(aligned16(part.p[g]) && aligned16(dst.p[g]), "tmf_peer_reduce_push: buffers must be 16-byte aligned");
  const long long n4 = n_elems / 4, off4 = elem_off / 4;
  const int grid = (int)std::min<long long>(cdiv(n4, 256), 8ll * kNumSMs);
  cudaStream_t st = as_stream(stream);
  if (world <= 2) peer_reduce_push_kernel<2><<<grid, 256, 0, st>>>(part, dst, off4, n4, world, rank, lr);
  else if (world <= 4) peer_reduce_push_kernel<4><<<grid, 256, 0, st>>>(part, dst, off4, n4, world, rank, lr);
  else if (world <= 8) peer_reduce_push_kernel<8><<<grid, 256, 0, st>>>(part, dst, off4, n4, world, rank, lr);
  else peer_reduce_push_kernel<16><<<grid, 256, 0, st>>>(part, dst, off4, n4, world, rank, lr);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}
