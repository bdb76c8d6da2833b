// Fused U.V^T + per-row top-k for sm_100a:  tcgen05 bf16 GEMM (TMA-staged operands, fp32 accumulators
// in TMEM) whose epilogue never writes scores to HBM -- it filters each accumulator tile against a
// per-row running threshold and appends the few survivors to a candidate list; an fp64-accumulated
// rerank of the candidates then makes the returned indices exact (ties -> lower item id).
//
// Exactness argument (DESIGN.md "top-k"): the bf16 GEMM score s~ of a pair differs from the canonical
// score s by at most e = 2^-8 * 1.05 * |u| * |v|.  With t~ the k-th largest s~ seen so far in a row,
// every item of the final top-k satisfies s~ >= t~ - 2E (E = e with |v| := max |v|), so the candidate
// list is a superset of the answer; the rerank sorts it by (canonical score desc, item id asc).
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace tmf {

constexpr int BM = 128;        // users per CTA tile (TMEM lanes)
constexpr int BN = 128;        // items per accumulator tile (TMEM columns); two CTAs share an SM
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NSTAGES = 4;     // B-operand ring (16 KB stages)
constexpr int TMEM_COLS = 2 * BN;  // two accumulators per CTA, double-buffered against the epilogue
constexpr int CAP = 512;       // candidate slots per row
constexpr int CPL = CAP / 32;  // candidates per lane in warp-cooperative passes
constexpr int TOPK_THREADS = 256;
constexpr int A_SUB_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2; // 16 KB
constexpr int MAX_KB = 4;                  // n_components <= 256
constexpr float ERR_FACTOR = 1.05f / 256.0f;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait that also pins the loaded registers behind it (consumers cannot be scheduled above the wait)
__device__ __forceinline__ void tmem_ld_wait_for(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);  // start address
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------ warp-cooperative selection
__device__ __forceinline__ uint32_t f2key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// k-th largest of the n (<= CAP) values spread 16-per-lane (element e = lane + 32 t); hist: 256 ints of smem
__device__ float warp_kth_largest(const float (&sc)[CPL], int n, int k, int* hist) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0, mask = 0;
  int krem = k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const uint32_t key = f2key(sc[t]);
      if (lane + 32 * t < n && (key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncwarp();
    int c[8];
    int lsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // lane L owns bins 255-8L .. 248-8L, visited in descending order
      c[i] = hist[255 - 8 * lane - i];
      lsum += c[i];
    }
    int incl = lsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned reach = __ballot_sync(0xffffffffu, incl >= krem);
    const int F = reach ? __ffs(reach) - 1 : 31;  // reach != 0 whenever n >= k
    int bin = 0, knew = 0;
    if (lane == F) {
      int cum = incl - lsum;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (cum + c[i] >= krem) { bin = 255 - 8 * lane - i; knew = krem - cum; break; }
        cum += c[i];
      }
    }
    bin = __shfl_sync(0xffffffffu, bin, F);
    krem = __shfl_sync(0xffffffffu, knew, F);
    prefix |= (uint32_t)bin << shift;
    mask |= 255u << shift;
    __syncwarp();
  }
  return key2f(prefix);
}

struct TopkParams {
  long long n_users, n_items;   // real sizes
  int n_ublocks, n_tiles, kb;   // padded tiling: user blocks of 128, item tiles of 256, k-blocks of 64
  int k, clamp, item_offset;
  const float* unorm;           // [n_users_pad] l2 norm of each user row
  const float* vmax;            // [1] max item-row norm
  float2* cand;                 // [n_users_pad][CAP] (approx score, item id bits)
  int* cnt;                     // [n_users_pad] candidates per row, -1 = overflow (exact path)
  int* ovf_count;               // [1]
  int* ovf_rows;                // [n_users_pad]
  float* dump;                  // optional [n_users][dump_ld]: raw bf16-GEMM scores (bring-up / error-bound tests)
  long long dump_ld;
};

// keep rule shared by the in-kernel compaction and the final selection
__device__ __forceinline__ float new_threshold(float kth, float E, int clamp) {
  return clamp ? fmaxf(kth - E, 0.f) - E : kth - 2.f * E;
}

// compact one row's candidate list in place; returns the new count and threshold (all lanes)
__device__ void warp_compact(float2* buf, int n, int k, float E, int clamp, int item_offset, int* hist, int& n_out, float& thr_out) {
  const int lane = threadIdx.x & 31;
  float sc[CPL];
  int ix[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int e = lane + 32 * t;
    float2 x = make_float2(-INFINITY, 0.f);
    if (e < n) x = buf[e];
    sc[t] = x.x;
    ix[t] = __float_as_int(x.y);
  }
  const float kth = warp_kth_largest(sc, n, k, hist);
  const float thr = new_threshold(kth, E, clamp);
  __syncwarp();
  int base = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int e = lane + 32 * t;
    const bool keep = e < n && (sc[t] >= thr || (clamp && ix[t] - item_offset < k));
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(b & lt)] = make_float2(sc[t], __int_as_float(ix[t]));
    base += __popc(b);
  }
  __syncwarp();
  n_out = base;
  thr_out = thr;
}

// Filter one 32-column slice of a row's accumulator against its running threshold.  The fast path is a
// 3-input max tree (FMNMX3) and one compare; 8-column groups that contain a survivor are rescanned.
__device__ __forceinline__ void epilogue_chunk(uint32_t (&r)[32], int col0, bool tail_tile, bool valid, float thr, int& cnt,
                                               float2* buf, long long row, const TopkParams& p) {
  if (tail_tile) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= p.n_items) r[j] = 0xff800000u;  // -inf: padded items never qualify
  }
  if (p.dump != nullptr && valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < p.n_items) p.dump[row * p.dump_ld + col0 + j] = __uint_as_float(r[j]);
  }
  float m8[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float a = fmaxf(fmaxf(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
    const float b = fmaxf(fmaxf(a, __uint_as_float(r[8 * g + 3])), __uint_as_float(r[8 * g + 4]));
    const float c = fmaxf(fmaxf(b, __uint_as_float(r[8 * g + 5])), __uint_as_float(r[8 * g + 6]));
    m8[g] = fmaxf(c, __uint_as_float(r[8 * g + 7]));
  }
  const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  const bool filler = p.clamp && col0 < p.k;  // clamp mode: the k lowest item ids of the slab are always kept
  if (!__any_sync(0xffffffffu, valid && (mx >= thr || filler))) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const bool gh = valid && (m8[g] >= thr || (filler && col0 + 8 * g < p.k));
    if (__any_sync(0xffffffffu, gh)) {
      if (gh) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = __uint_as_float(r[8 * g + j]);
          const int col = col0 + 8 * g + j;
          if (col < p.n_items && (v >= thr || (filler && col < p.k))) {
            buf[cnt] = make_float2(v, __int_as_float(p.item_offset + col));
            ++cnt;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the fused kernel
__global__ void __launch_bounds__(TOPK_THREADS, 2)
score_topk_kernel(const __grid_constant__ CUtensorMap tmapU, const __grid_constant__ CUtensorMap tmapV, const TopkParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve (1024-byte aligned operand tiles first)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // kb sub-tiles of [128][64] bf16
  unsigned char* sB = sA + p.kb * A_SUB_BYTES;                // NSTAGES x [256][64] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGES * B_STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [NSTAGES]
  uint64_t* empty_bar = bars + NSTAGES;       // [NSTAGES]
  uint64_t* a_full = bars + 2 * NSTAGES;      // [1]
  uint64_t* a_empty = a_full + 1;             // [1]
  uint64_t* tfull = a_empty + 1;              // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* hist_all = reinterpret_cast<int*>(tmem_slot + 4);  // 4 warps x 256

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapU) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapV) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(a_full), 1);
    mbar_init(smem_u32(a_empty), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: 256 columns = two 128x128 fp32 accumulators (the SM's other CTA takes the other half)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
        mbar_wait(smem_u32(a_empty), a_phase ^ 1);  // previous user block's MMAs retired
        mbar_expect_tx(smem_u32(a_full), p.kb * A_SUB_BYTES);
        for (int kb = 0; kb < p.kb; ++kb) tma_load_2d(smem_u32(sA + kb * A_SUB_BYTES), &tmapU, kb * BK, ub * BM, smem_u32(a_full));
        a_phase ^= 1;
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            mbar_expect_tx(smem_u32(&full_bar[stage]), B_STAGE_BYTES);
            tma_load_2d(smem_u32(sB + stage * B_STAGE_BYTES), &tmapV, kb * BK, nt * BN, smem_u32(&full_bar[stage]));
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
        mbar_wait(smem_u32(a_full), a_phase);
        a_phase ^= 1;
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          mbar_wait(smem_u32(&tempty[acc]), acc_phase ^ 1);  // epilogue drained this accumulator
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(sA + kb * A_SUB_BYTES);
            const uint32_t b_addr = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              tcgen05_mma_f16(d_tmem, umma_desc_sw128(a_addr + k * UMMA_K * 2), umma_desc_sw128(b_addr + k * UMMA_K * 2), kIdesc,
                              (uint32_t)((kb | k) != 0));
            }
            tcgen05_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot when these MMAs retire
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
          tcgen05_commit(smem_u32(&tfull[acc]));  // accumulator ready for the epilogue
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        tcgen05_commit(smem_u32(a_empty));  // A tile may be overwritten
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: threshold filter + candidate append =====================
    const int q = warp - 4;  // TMEM lane quarter == warp % 4
    int* hist = hist_all + q * 256;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float vmax = *p.vmax;
    for (int ub = blockIdx.x; ub < p.n_ublocks; ub += gridDim.x) {
      const long long row = (long long)ub * BM + q * 32 + lane;
      const bool valid = row < p.n_users;
      const float E = ERR_FACTOR * p.unorm[row] * vmax + 1e-30f;
      float thr = valid ? -INFINITY : INFINITY;
      int cnt = 0;
      float2* buf = p.cand + row * CAP;
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        mbar_wait(smem_u32(&tfull[acc]), acc_phase);
        __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the per-lane spin
        tcgen05_fence_after();
        const bool tail_tile = (long long)(nt + 1) * BN > p.n_items;
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
        uint32_t ra[32], rb[32];
        tmem_ld32(t_base, ra);
#pragma unroll
        for (int ch = 0; ch < BN / 32; ch += 2) {
          // chunk ch is in ra; chunk ch+1 is fetched into rb while ra is filtered (and vice versa)
          tmem_ld_wait_for(ra);
          tmem_ld32(t_base + (uint32_t)((ch + 1) * 32), rb);
          epilogue_chunk(ra, nt * BN + ch * 32, tail_tile, valid, thr, cnt, buf, row, p);
          tmem_ld_wait_for(rb);
          if (ch + 2 < BN / 32) tmem_ld32(t_base + (uint32_t)((ch + 2) * 32), ra);
          epilogue_chunk(rb, nt * BN + (ch + 1) * 32, tail_tile, valid, thr, cnt, buf, row, p);
        }
        // accumulator drained: hand it back to the MMA warp before any list maintenance
        tcgen05_fence_before();
        mbar_arrive(smem_u32(&tempty[acc]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        // make room for the next tile's (at most BN) appends -- warp-uniform
        unsigned need = __ballot_sync(0xffffffffu, cnt > CAP - BN);
        while (need) {
          const int owner = __ffs(need) - 1;
          need &= need - 1;
          const int n_o = __shfl_sync(0xffffffffu, cnt, owner);
          const float E_o = __shfl_sync(0xffffffffu, E, owner);
          float2* buf_o = p.cand + ((long long)ub * BM + q * 32 + owner) * CAP;
          __syncwarp();
          int n_new;
          float thr_new;
          warp_compact(buf_o, n_o, p.k, E_o, p.clamp, p.item_offset, hist, n_new, thr_new);
          if (lane == owner) {
            if (n_new > CAP - BN) {  // cannot shrink (massive ties): hand the row to the exact path
              const int slot = atomicAdd(p.ovf_count, 1);
              p.ovf_rows[slot] = (int)row;
              thr = INFINITY;
              cnt = -1;
            } else {
              thr = thr_new;
              cnt = n_new;
            }
          }
        }
      }
      p.cnt[row] = valid ? cnt : 0;
    }
  }

  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ final selection + canonical rerank
constexpr int RR_WARPS = 4;

struct RerankParams {
  long long n_users, n_items;
  int k, clamp, item_offset, r, ld;
  const float* U;
  const float* V;
  const float* unorm;
  const float* vmax;
  const float2* cand;
  const int* cnt;
  int* out_idx;
  float* out_score;
};

__global__ void __launch_bounds__(RR_WARPS * 32) rerank_kernel(const RerankParams p) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rp = p.r + 1;  // padded row stride: conflict-free per-lane row walks
  const size_t per_warp = (size_t)(32 * rp + p.r) * sizeof(float) + CAP * 8 + 256 * sizeof(int);
  unsigned char* base = rr_smem + (size_t)w * per_warp;
  float* tile = reinterpret_cast<float*>(base);      // [32][r+1] gathered item rows
  float* su = tile + 32 * rp;                        // [r] this user's row
  float* ex = su + p.r;                              // [CAP] exact scores
  int* id = reinterpret_cast<int*>(ex + CAP);        // [CAP] item ids
  int* hist = id + CAP;                              // [256]

  const long long row = (long long)blockIdx.x * RR_WARPS + w;
  if (row >= p.n_users) return;
  const int n = p.cnt[row];
  if (n < 0) return;  // overflowed row: exact_rows_kernel owns it
  const float2* buf = p.cand + row * CAP;
  float sc[CPL];
  int ix[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int e = lane + 32 * t;
    float2 x = make_float2(-INFINITY, 0.f);
    if (e < n) x = buf[e];
    sc[t] = x.x;
    ix[t] = __float_as_int(x.y);
  }
  const int k = p.k;
  const float E = ERR_FACTOR * p.unorm[row] * (*p.vmax) + 1e-30f;
  const float kth = warp_kth_largest(sc, n, k, hist);
  const float thr = new_threshold(kth, E, p.clamp);
  int m = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int e = lane + 32 * t;
    const bool keep = e < n && (sc[t] >= thr || (p.clamp && ix[t] - p.item_offset < k));
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) id[m + __popc(b & lt)] = ix[t];
    m += __popc(b);
  }
  for (int c = lane; c < p.r; c += 32) su[c] = p.U[row * p.ld + c];
  __syncwarp();
  // canonical scores, 32 candidates at a time: coalesced row gathers into smem, then one fp64 FMA
  // chain per lane in component order (bit-identical to oracle.canonical_scores)
  for (int b0 = 0; b0 < m; b0 += 32) {
    const int nb = min(32, m - b0);
    for (int j = 0; j < nb; ++j) {
      const float* vrow = p.V + (long long)(id[b0 + j] - p.item_offset) * p.ld;
      for (int c = lane; c < p.r; c += 32) tile[j * rp + c] = vrow[c];
    }
    __syncwarp();
    if (lane < nb) {
      double acc = 0.0;
      const float* tr = tile + lane * rp;
      for (int c = 0; c < p.r; ++c) acc = fma((double)su[c], (double)tr[c], acc);
      float s = (float)acc;
      if (p.clamp) s = s > 0.f ? s : 0.f;  // tf.where(p > 0, p, 0.0)
      ex[b0 + lane] = s + 0.0f;
    }
    __syncwarp();
  }
  // rank by counting with comparator (score desc, item id asc); ids are distinct so ranks are too
  for (int t = lane; t < m; t += 32) {
    const float s = ex[t];
    const int my = id[t];
    int rank = 0;
    for (int o = 0; o < m; ++o) rank += (ex[o] > s) || (ex[o] == s && id[o] < my);
    if (rank < k) {
      p.out_idx[row * k + rank] = my;
      p.out_score[row * k + rank] = s;
    }
  }
}

// exact path for rows whose candidate list overflowed (huge tie groups): all canonical scores of the
// row into scratch, then k rounds of block-wide arg-max in (score desc, id asc) order.
__global__ void __launch_bounds__(256) exact_rows_kernel(const RerankParams p, const int* __restrict__ ovf_count,
                                                         const int* __restrict__ ovf_rows, float* __restrict__ scratch) {
  __shared__ float s_best[8];
  __shared__ int s_besti[8];
  __shared__ float s_prev;
  __shared__ int s_previ;
  const int n_ovf = *ovf_count;
  float* sc = scratch + (long long)blockIdx.x * p.n_items;
  for (int o = blockIdx.x; o < n_ovf; o += gridDim.x) {
    const long long row = ovf_rows[o];
    const float* u = p.U + row * p.ld;
    for (long long i = threadIdx.x; i < p.n_items; i += 256) {
      const float* v = p.V + i * p.ld;
      double acc = 0.0;
      for (int c = 0; c < p.r; ++c) acc = fma((double)u[c], (double)v[c], acc);
      float s = (float)acc;
      if (p.clamp) s = s > 0.f ? s : 0.f;
      sc[i] = s + 0.0f;
    }
    if (threadIdx.x == 0) { s_prev = INFINITY; s_previ = -1; }
    __syncthreads();
    for (int q = 0; q < p.k; ++q) {
      const float ps = s_prev;
      const int pi = s_previ;
      float best = -INFINITY;
      int besti = 0x7fffffff;
      for (long long i = threadIdx.x; i < p.n_items; i += 256) {
        const float s = sc[i];
        const bool after = (s < ps) || (s == ps && (int)i > pi);
        if (after && (s > best || (s == best && (int)i < besti))) { best = s; besti = (int)i; }
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o2);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
      }
      if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_besti[threadIdx.x >> 5] = besti; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w2 = 1; w2 < 8; ++w2)
          if (s_best[w2] > best || (s_best[w2] == best && s_besti[w2] < besti)) { best = s_best[w2]; besti = s_besti[w2]; }
        p.out_idx[row * p.k + q] = besti + p.item_offset;
        p.out_score[row * p.k + q] = best;
        s_prev = best;
        s_previ = besti;
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ operand packing
// one warp per row: fp32 [n, ld] -> bf16 [n_pad, k_pad] (zero padded), row norm, optional global max norm
__global__ void pack_bf16_kernel(const float* __restrict__ src, long long n, int r, int ld, __nv_bfloat16* __restrict__ dst,
                                 long long n_pad, int k_pad, float* __restrict__ norms, int* __restrict__ max_norm_bits) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  float ss = 0.f;
  for (int c = lane; c < k_pad; c += 32) {
    float v = 0.f;
    if (row < n && c < r) v = src[row * ld + c];
    ss = fmaf(v, v, ss);
    dst[row * k_pad + c] = __float2bfloat16_rn(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) {
    const float nrm = sqrtf(ss) * 1.0001f;
    norms[row] = nrm;
    if (max_norm_bits) atomicMax(max_norm_bits, __float_as_int(nrm));  // non-negative floats order like ints
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int make_tmap(CUtensorMap* map, void* base, long long rows, int k_pad, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  TMF_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TMF_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return TMF_OK;
}

struct TopkLayout {
  long long nu_pad, ni_pad;
  int k_pad, kb;
  size_t off_ub, off_vb, off_unorm, off_vnorm, off_vmax, off_cand, off_cnt, off_ovfc, off_ovfr, off_scratch, total;
  int scratch_rows;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static TopkLayout topk_layout(long long n_users, long long n_items, int r) {
  TopkLayout L{};
  L.nu_pad = cdiv(n_users, BM) * BM;
  L.ni_pad = cdiv(n_items, BN) * BN;
  L.k_pad = (int)(cdiv(r, BK) * BK);
  L.kb = L.k_pad / BK;
  size_t o = 0;
  L.off_ub = o; o = align_up(o + (size_t)L.nu_pad * L.k_pad * 2, 1024);
  L.off_vb = o; o = align_up(o + (size_t)L.ni_pad * L.k_pad * 2, 1024);
  L.off_unorm = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.off_vnorm = o; o = align_up(o + (size_t)L.ni_pad * 4, 256);
  L.off_vmax = o; o += 256;
  L.off_cand = o; o = align_up(o + (size_t)L.nu_pad * CAP * 8, 256);
  L.off_cnt = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.off_ovfc = o; o += 256;
  L.off_ovfr = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.scratch_rows = (int)std::min<long long>(32, n_users);
  L.off_scratch = o; o = align_up(o + (size_t)L.scratch_rows * n_items * 4, 256);
  L.total = o + 1024;
  return L;
}

}  // namespace tmf

using namespace tmf;

extern "C" int tmf_pack_bf16(const float* src, int64_t n, int32_t n_comp, int32_t ld, uint16_t* dst, int64_t n_pad,
                             int32_t k_pad, float* norms, tmf_stream_t stream) {
  TMF_REQUIRE(n_pad >= n && k_pad >= n_comp && n_comp <= ld, "tmf_pack_bf16: bad shape");
  if (n_pad == 0) return TMF_OK;
  pack_bf16_kernel<<<(unsigned)cdiv(n_pad * 32, 256), 256, 0, as_stream(stream)>>>(src, n, n_comp, ld, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                                n_pad, k_pad, norms, nullptr);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}

extern "C" size_t tmf_score_topk_ws_bytes(int64_t n_users, int64_t n_items, int32_t n_comp, int32_t k) {
  (void)k;
  return topk_layout(n_users, n_items, n_comp).total;
}

static int score_topk_impl(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                           int32_t k, int32_t clamp, int32_t item_offset, int32_t* out_idx, float* out_score, void* ws,
                           size_t ws_bytes, tmf_stream_t stream, float* dump) {
  TMF_REQUIRE(n_users >= 0 && n_items > 0 && n_comp > 0 && n_comp <= ld, "tmf_score_topk: bad shape");
  TMF_REQUIRE(n_comp <= MAX_KB * BK, "tmf_score_topk: n_components up to %d supported", MAX_KB * BK);
  TMF_REQUIRE(k >= 1 && k <= 128 && k <= n_items, "tmf_score_topk: need 1 <= k <= min(128, n_items) (k=%d)", k);
  TMF_REQUIRE(n_items < (1ll << 31) - BN && n_users < (1ll << 31) - BM, "tmf_score_topk: sizes must fit int32");
  if (n_users == 0) return TMF_OK;
  const TopkLayout L = topk_layout(n_users, n_items, n_comp);
  TMF_REQUIRE(ws_bytes >= L.total, "tmf_score_topk: workspace too small");
  unsigned char* w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* Ub = reinterpret_cast<__nv_bfloat16*>(w + L.off_ub);
  __nv_bfloat16* Vb = reinterpret_cast<__nv_bfloat16*>(w + L.off_vb);
  float* unorm = reinterpret_cast<float*>(w + L.off_unorm);
  float* vnorm = reinterpret_cast<float*>(w + L.off_vnorm);
  int* vmax_bits = reinterpret_cast<int*>(w + L.off_vmax);
  float2* cand = reinterpret_cast<float2*>(w + L.off_cand);
  int* cnt = reinterpret_cast<int*>(w + L.off_cnt);
  int* ovfc = reinterpret_cast<int*>(w + L.off_ovfc);
  int* ovfr = reinterpret_cast<int*>(w + L.off_ovfr);
  float* scratch = reinterpret_cast<float*>(w + L.off_scratch);
  cudaStream_t st = as_stream(stream);

  TMF_CUDA(cudaMemsetAsync(vmax_bits, 0, 4, st));
  TMF_CUDA(cudaMemsetAsync(ovfc, 0, 4, st));
  pack_bf16_kernel<<<(unsigned)cdiv(L.nu_pad * 32, 256), 256, 0, st>>>(U, n_users, n_comp, ld, Ub, L.nu_pad, L.k_pad, unorm, nullptr);
  pack_bf16_kernel<<<(unsigned)cdiv(L.ni_pad * 32, 256), 256, 0, st>>>(V, n_items, n_comp, ld, Vb, L.ni_pad, L.k_pad, vnorm, vmax_bits);
  TMF_LAUNCH_CHECK();

  CUtensorMap tmapU, tmapV;
  int rc = make_tmap(&tmapU, Ub, L.nu_pad, L.k_pad, BM);
  if (rc) return rc;
  rc = make_tmap(&tmapV, Vb, L.ni_pad, L.k_pad, BN);
  if (rc) return rc;

  TopkParams p{};
  p.n_users = n_users; p.n_items = n_items;
  p.n_ublocks = (int)(L.nu_pad / BM); p.n_tiles = (int)(L.ni_pad / BN); p.kb = L.kb;
  p.k = k; p.clamp = clamp ? 1 : 0; p.item_offset = item_offset;
  p.unorm = unorm; p.vmax = reinterpret_cast<const float*>(vmax_bits);
  p.cand = cand; p.cnt = cnt; p.ovf_count = ovfc; p.ovf_rows = ovfr;
  p.dump = dump; p.dump_ld = n_items;

  const size_t smem = 1024 + (size_t)L.kb * A_SUB_BYTES + (size_t)NSTAGES * B_STAGE_BYTES + 256 + 4 * 256 * sizeof(int);
  TMF_CUDA(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(2 * kNumSMs, p.n_ublocks);  // two CTAs per SM (smem- and TMEM-limited)
  score_topk_kernel<<<grid, TOPK_THREADS, smem, st>>>(tmapU, tmapV, p);
  TMF_LAUNCH_CHECK();
  if (dump != nullptr) return TMF_OK;

  RerankParams q{};
  q.n_users = n_users; q.n_items = n_items; q.k = k; q.clamp = p.clamp; q.item_offset = item_offset; q.r = n_comp; q.ld = ld;
  q.U = U; q.V = V; q.unorm = unorm; q.vmax = p.vmax; q.cand = cand; q.cnt = cnt; q.out_idx = out_idx; q.out_score = out_score;
  const size_t rr_smem = (size_t)RR_WARPS * ((size_t)(32 * (n_comp + 1) + n_comp) * sizeof(float) + CAP * 8 + 256 * sizeof(int));
  TMF_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
  rerank_kernel<<<(unsigned)cdiv(n_users, RR_WARPS), RR_WARPS * 32, rr_smem, st>>>(q);
  TMF_LAUNCH_CHECK();
  exact_rows_kernel<<<L.scratch_rows, 256, 0, st>>>(q, ovfc, ovfr, scratch);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}

extern "C" int tmf_score_topk(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                              int32_t k, int32_t clamp, int32_t item_offset, int32_t* out_idx, float* out_score, void* ws,
                              size_t ws_bytes, tmf_stream_t stream) {
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, k, clamp, item_offset, out_idx, out_score, ws, ws_bytes, stream, nullptr);
}

extern "C" int tmf_score_dense_bf16(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                                    float* P, void* ws, size_t ws_bytes, tmf_stream_t stream) {
  TMF_REQUIRE(P != nullptr, "tmf_score_dense_bf16: null output");
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, 1, 0, 0, nullptr, nullptr, ws, ws_bytes, stream, P);
}
