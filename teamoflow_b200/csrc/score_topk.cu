// Fused U.V^T + per-row top-k for sm_100a:  tcgen05 bf16 GEMM (TMA-staged operands, fp32 accumulators
// in TMEM) whose epilogue never writes scores to HBM -- it filters each accumulator tile against a
// per-row running threshold and appends the few survivors to a candidate list; an fp64-accumulated
// rerank of the candidates then makes the returned indices exact (ties -> lower item id).
//
// Exactness argument (DESIGN.md "top-k"): with u~ = fl16(u), du = u - u~ (both known exactly) the tensor-core
// score s~ of a pair differs from the canonical score s by at most E = |du| max|v| + |u~| max|dv| + accumulation
// slack (row_error_kernel; rigorous, computed from the data -- the worst case for bf16 is 2^-7 |u||v|).  With t~ a
// lower bound of the k-th largest s~ seen so far in a row, every item of the final top-k satisfies s~ >= t~ - 2E,
// so the candidate list is a superset of the answer; the rerank sorts it by (canonical score desc, item id asc).
//
// Developer switches (TMF_TOPK_DEBUG / TMF_TOPK_FMT / TMF_TOPK_PROF environment variables) exist only in builds
// compiled with -DTMF_DEVTOOLS (scripts/build_variant.sh); the product library reads no environment variable.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmf {

constexpr int BM = 128;        // users per CTA tile (TMEM lanes)
constexpr int BN = 128;        // items per MMA tile = per accumulator buffer (TMEM columns)
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 12; // B-operand ring (32 KB stages, 16 KB per CTA of a pair); the launch uses as many as shared memory allows (>= 2)
// ONE CTA per SM owns all 512 TMEM columns = two 128 x 256 fp32 accumulator buffers, and runs four epilogue GROUPS of four warps
// (16 epilogue warps per SM, four per sub-partition, where round 1's two-CTA layout had eight).  The MMAs are N = 256 wide:
// at N = 128 every MMA reads 4 KB of A and 4 KB of B from shared memory per 64 cycles -- the whole 128 B/clk of the SM's shared
// memory, before the TMA writes -- and costs the issuing thread as much as an N = 256 one.
constexpr int NBUF = 4;         // accumulator buffers of BN columns (all 512 TMEM columns); buffer b holds the tiles nt % 4 == b
constexpr int NGRP = NBUF;      // epilogue groups (4 warps each): group g owns buffer g, i.e. filters the tiles nt % 4 == g
constexpr int QN = 64;          // columns per filter pass (four x16 TMEM loads, one wait)
constexpr int TMEM_COLS = NBUF * BN;
// A row's candidates are appended per group (group g sees every fourth tile of the row): CAPG slots each.  The running
// threshold is per ROW: one score histogram per row in shared memory, fed by all four groups (red.shared), from which any
// group derives "the highest bin edge with >= k entries at or above it" -- a lower bound of the row's k-th best score so far.
// (A first version kept a private histogram per group: each group then tracks the k-th best of ITS quarter of the items,
// roughly the 4k-th best overall, 3.4x the survivors and 525 ms where the two-CTA kernel took 322.)
constexpr int CAPG = 1024;
constexpr int CAP = NGRP * CAPG;  // candidate slots per row in the workspace: [row][group][CAPG]
constexpr int NBINS = 48;      // per-row score histogram bins (32-bit counts)
#ifndef TMF_QCAP
#define TMF_QCAP 16
#endif
#ifndef TMF_PF_TILES
#define TMF_PF_TILES 0
#endif
constexpr int QCAP = TMF_QCAP;       // per-(row, group) survivor queue slots in shared memory (drained warp-wide)
constexpr int HSTRIDE = NBINS + 1;  // words per row: odd, so the lanes' rows fall into different banks
constexpr int UB_BATCH = 2048; // user blocks (x128 rows) per main-kernel launch: bounds the candidate workspace to 8.6 GB
constexpr int TOPK_THREADS = 128 + NGRP * 128;  // TMA, MMA, TMEM-alloc, (idle) warps + 4 groups x 4 epilogue warps
constexpr int A_SUB_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2; // 16 KB
constexpr int MAX_KB = 4;                  // n_components <= 256
// Error bound of the tensor-core score s~ = fl32(sum u~_c v~_c) against the real score s = sum u_c v_c, with u~ = bf16(u),
// du = u - u~ (both known exactly):  s - sum u~_c v~_c = du.v + u~.dv  =>  |s~ - s| <= |du||v| + |u~||dv| + accumulation.
// The residual norms are computed from the data by pack_bf16_kernel, so the bound is rigorous (unit roundoff of bf16 is
// 2^-8 per operand: worst case 2^-7 |u||v|) AND tight for typical data (about 0.8 * 2^-8 |u||v|).
// ACC_UNIT: slack per accumulated k-step for the fp32 accumulation inside the tensor core, relative to |u~||v~|.
constexpr float ACC_UNIT = 2.4e-7f;
constexpr float NORM_SLACK = 1.0001f;  // rounding of the fp32 norm arithmetic itself

#ifdef TMF_DEVTOOLS
#define TMF_DBG(p) ((p).dbg)
#else
#define TMF_DBG(p) 0
#endif

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

// shared-space reduction (a generic-address atomicAdd compiles to the slow generic ATOM path)
__device__ __forceinline__ void smem_inc(int* p) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

__device__ __forceinline__ bool use_fp16(const float* stats);  // operand format choice, defined with the pack kernels

struct TopkParams {
  long long n_users, n_items;   // real sizes
  int ub0, n_ublocks;           // this launch covers user blocks [ub0, ub0 + n_ublocks)
  int n_tiles, kb;              // item tiles of BN, k-blocks of BK
  int nstages;                  // B-operand ring depth of this launch
  int k, clamp, item_offset;
  const float* erow;            // [n_users_pad] error bound E of each user row against this item slab (global row index)
  float2* cand;                 // [batch rows][NGRP][CAPG] (approx score, item id bits), batch-local row index
  int* cnt;                     // [batch rows][NGRP] candidates per (row, group), -1 = overflow (the rerank hands the row to the exact path)
  float* thr_out;               // [batch rows][NGRP] final keep-threshold of each group (each is valid for the whole row)
  float* dump;                  // optional [n_users][dump_ld]: raw bf16-GEMM scores (bring-up / error-bound tests)
  long long dump_ld;
  const float* fmt_stats;       // operand statistics (see use_fp16); NULL with force_fmt >= 0
  int force_fmt;                // -1 = by fmt_stats, 0 = bf16, 1 = fp16
  const float* row_bound;       // optional [n_users]: lower bounds of the rows' k-th best TRUE score (cross-GPU exchange), raw mode only
  unsigned long long* prof;     // optional [8] cycle counters (env TMF_TOPK_PROF=1): where the warps wait
  int dbg;                      // profiling aid (env TMF_TOPK_DEBUG): 1 = no appends, 2 = no filtering, 3 = no TMEM reads
};

// Keep-threshold from a lower bound `kth` of the row's k-th largest approximate score.
//   raw scores   : every top-k member has s~ >= kth - 2E
//   clamped (>0) : every top-k member outside the k lowest item ids has s~ >= max(kth - E, 0) - E
__device__ __forceinline__ float keep_threshold(float kth, float E, int clamp) {
  return clamp ? fmaxf(kth - E, 0.f) - E : kth - 2.f * E;
}

// ---- per-row running threshold: a 48-bin histogram of the appended scores per ROW in shared memory (32-bit counts),
// shared by the row's four epilogue groups and double-buffered by user-block parity.  "The highest bin whose at-or-above
// count is >= k" gives a lower bound of the k-th largest score seen so far: a scan that races with other groups'
// increments can only UNDER-count, so every threshold it yields is valid; the published threshold (ordered uint32 key,
// red.shared.max) only rises.  The bin range is set once per row from the row's first tile: [mean, mean + 2.5 (max - mean)).
//
// ---- SIMD list maintenance.  A lane that finds a survivor only pushes (score, item) onto its own small
// shared-memory queue (3 instructions, no dependent loads).  When a queue is about to fill, the WHOLE warp
// drains: iteration s handles entry s of every lane at once, each lane appending to its own (row, group) list and
// counting the entry in its row's histogram -- the append code runs once per queue slot for 32 rows instead of
// once per survivor with 1-3 active lanes (measured: 400-600 cycles per survivor in the divergent versions).
struct RowState {
  float thr;      // current keep-threshold (+inf for padded rows and for lists handed to the exact path)
  float thr_ext;  // externally supplied floor of the threshold (-inf when there is none)
  float lo, w, inv_w, E;
  int cnt;        // list length; > CAPG = saturated (-> exact path)
  int cq;         // entries waiting in the lane's queue
};

// explicit shared-space accesses: through generic pointers these compile to the slower ST.E/LD.E with 64-bit addressing
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float2 lds_f2(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }

__device__ __forceinline__ float edge_threshold(float lo, float w, int bthr, float E, int clamp) {
  const float edge = lo + (float)bthr * w;
  const float slack = 4e-7f * (fabsf(lo) + (float)NBINS * w);  // rounding of the bin arithmetic
  return keep_threshold(edge - slack, E, clamp);
}

// all lanes, each for its own row: highest bin with >= k entries at or above it (-1 when the row has fewer than k binned
// entries).  Walks down from the top bin; the warp stops when every lane has its answer.
__device__ __forceinline__ int hist_threshold_bin(uint32_t hrow, int k) {
  static_assert(NBINS % 16 == 0, "scanned in chunks of 16 bins");
  int cum = 0, found = -1;
#pragma unroll 1
  for (int c = NBINS / 16 - 1; c >= 0; --c) {
    uint32_t v[16];  // 16 independent loads in flight (a bin-by-bin walk pays the shared-memory latency 48 times: ~2500 cycles)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = lds_u32_volatile(hrow + 4u * (uint32_t)(c * 16 + j));
#pragma unroll
    for (int j = 15; j >= 0; --j) {
      cum += (int)v[j];
      if (found < 0 && cum >= k) found = c * 16 + j;
    }
    if (__all_sync(0xffffffffu, found >= 0)) break;
  }
  return found;
}

// warp-wide drain of the per-lane queues (all lanes must call; cq may differ per lane): appends the queued survivors that
// still pass the row's threshold to the lane's list and counts them in the row's histogram (fire-and-forget red.shared),
// then re-derives the row's threshold from the histogram and publishes it.  Returns the new list length.
// (out of line, state by value in registers: inlining it at every call site costs registers in the tile loop)
struct DrainRet { float thr; int cnt; };
__device__ __noinline__ DrainRet drain_queues_nl(float thr, float thr_ext, float lo, float w, float inv_w, float E, int cnt, int cq,
                                                 uint32_t queue, float2* buf, uint32_t hrow, uint32_t thr_slot, int k, int clamp) {
  int maxq = cq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxq = max(maxq, __shfl_xor_sync(0xffffffffu, maxq, o));
#pragma unroll 4
  for (int s = 0; s < maxq; ++s) {
    const float2 e = lds_f2(queue + 8u * (uint32_t)s);
    const bool ok = (s < cq) && (e.x >= thr);
    const bool okh = ok && (e.x >= lo);
    const int b = (int)fminf(fmaxf((e.x - lo) * inv_w, 0.f), (float)(NBINS - 1));
    const int st_ok = ok && (cnt < CAPG);
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %0, 0;\n\t"
        "setp.ne.s32 q, %1, 0;\n\t"
        "@p st.global.cg.v2.f32 [%2], {%3, %4};\n\t"
        "@q red.shared.add.u32 [%5], 1;\n\t}"
        ::"r"(st_ok), "r"((int)okh), "l"(buf + cnt), "f"(e.x), "f"(e.y), "r"(hrow + 4u * (uint32_t)b)
        : "memory");
    cnt += ok ? 1 : 0;
  }
  __syncwarp();
  if (maxq > 0) {  // warp-uniform: something may have been counted (queues only fill once the row state is live: lo is finite)
    const int bthr = hist_threshold_bin(hrow, k);
    if (bthr >= 0) {
      const float t = fmaxf(edge_threshold(lo, w, bthr, E, clamp), thr_ext);
      if (t > thr && thr < INFINITY) {
        thr = t;
        asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(thr_slot), "r"(f2key(t)) : "memory");
      }
    }
  }
  DrainRet r;
  r.thr = thr; r.cnt = cnt;
  return r;
}
// Per-tile share of the list maintenance: every lane moves (at most) the newest entry of its queue to its list, branch-free.  In
// steady state a lane queues a survivor every ~10 tiles, so this keeps the queues near empty at a FIXED cost per tile -- the
// accumulator buffers are handed back only when all epilogue warps (of both CTAs of a pair) are through with them, so a warp that
// every ~50 tiles spends a whole tile period in a full drain stalls the MMA threads and every other warp with it.
__device__ __forceinline__ void pop_one(RowState& st, uint32_t queue, float2* buf, uint32_t hrow) {
  const bool has = st.cq > 0;
  const float2 e = lds_f2(queue + 8u * (uint32_t)max(st.cq - 1, 0));
  const bool ok = has && (e.x >= st.thr);
  const bool okh = ok && (e.x >= st.lo);
  const int b = (int)fminf(fmaxf((e.x - st.lo) * st.inv_w, 0.f), (float)(NBINS - 1));
  const int st_ok = ok && (st.cnt < CAPG);
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.s32 p, %0, 0;\n\t"
      "setp.ne.s32 q, %1, 0;\n\t"
      "@p st.global.cg.v2.f32 [%2], {%3, %4};\n\t"
      "@q red.shared.add.u32 [%5], 1;\n\t}"
      ::"r"(st_ok), "r"((int)okh), "l"(buf + st.cnt), "f"(e.x), "f"(e.y), "r"(hrow + 4u * (uint32_t)b)
      : "memory");
  st.cnt += ok ? 1 : 0;
  st.cq -= has ? 1 : 0;
}
// the row's threshold from its histogram (warp-collective), published when it rose
__device__ __noinline__ float refresh_threshold(float thr, float thr_ext, float lo, float w, float E, uint32_t hrow, uint32_t thr_slot, int k, int clamp) {
  const int bthr = hist_threshold_bin(hrow, k);
  if (bthr >= 0) {
    const float t = fmaxf(edge_threshold(lo, w, bthr, E, clamp), thr_ext);
    if (t > thr && thr < INFINITY) {
      thr = t;
      asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(thr_slot), "r"(f2key(t)) : "memory");
    }
  }
  return thr;
}
__device__ __forceinline__ void drain_queues(RowState& st, uint32_t queue, float2* buf, uint32_t hrow, uint32_t thr_slot, int k, int clamp) {
  const DrainRet r = drain_queues_nl(st.thr, st.thr_ext, st.lo, st.w, st.inv_w, st.E, st.cnt, st.cq, queue, buf, hrow, thr_slot, k, clamp);
  st.thr = r.thr; st.cnt = r.cnt;
  st.cq = 0;
}

// dense-score dump of one 32-column slice (bring-up / error-bound tests and tmf_score_dense_*)
__device__ __forceinline__ void dump_chunk(const uint32_t (&r)[32], int col0, int n_items, bool valid, long long row, const TopkParams& p) {
  if (valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < n_items) p.dump[row * p.dump_ld + col0 + j] = __uint_as_float(r[j]);
  }
}

// hit bits of the two 8-column groups of one 16-column chunk: group maximum (3-input max tree, no branches) >= threshold
__device__ __forceinline__ unsigned chunk_hits(const uint32_t (&r)[16], float thr) {
  unsigned hm = 0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const float a = fmaxf(fmaxf(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
    const float b = fmaxf(fmaxf(a, __uint_as_float(r[8 * g + 3])), __uint_as_float(r[8 * g + 4]));
    const float c = fmaxf(fmaxf(b, __uint_as_float(r[8 * g + 5])), __uint_as_float(r[8 * g + 6]));
    hm |= (fmaxf(c, __uint_as_float(r[8 * g + 7])) >= thr) ? (1u << g) : 0u;
  }
  return hm;
}

// Steady-state filter of one group's QUARTER of a tile (64 of the accumulator buffer's 256 columns).  Pass 1
// streams the 64 columns through registers once -- four x16 loads issued back to back, ONE wait -- and keeps only the 8
// group-maximum hit bits: no votes, no branches.  One REDUX.OR of the per-lane hit masks then names the 8-column groups in which
// ANY row of the warp has a survivor; only those are re-read from TMEM (x8) and their survivors queued with predicated stores,
// all lanes convergent.
__device__ __forceinline__ void tmem_ld_wait_for16x4(uint32_t (&a)[16], uint32_t (&b)[16], uint32_t (&c)[16], uint32_t (&d)[16]) {
  tmem_ld_wait_for16x2(a, b);   // the first wait::ld completes all four loads; the second pins c and d behind a wait as well
  tmem_ld_wait_for16x2(c, d);
}
__device__ __forceinline__ void epilogue_tile(uint32_t t_base, int col0, RowState& st, uint32_t queue, float2* buf, uint32_t hrow,
                                              uint32_t thr_slot, const TopkParams& p) {
  uint32_t ra[16], rb[16], rc[16], rd[16];
  const float thr = st.thr;  // invalid rows carry thr = +inf; NaN-padded columns never win a max or a compare
  tmem_ld16(t_base, ra);
  tmem_ld16(t_base + 16u, rb);
  tmem_ld16(t_base + 32u, rc);
  tmem_ld16(t_base + 48u, rd);
  tmem_ld_wait_for16x4(ra, rb, rc, rd);
  if (TMF_DBG(p) == 5) return;  // diagnostic: accumulator reads only
  const unsigned hm = chunk_hits(ra, thr) | (chunk_hits(rb, thr) << 2) | (chunk_hits(rc, thr) << 4) | (chunk_hits(rd, thr) << 6);
  unsigned gmask = __reduce_or_sync(0xffffffffu, hm);
  if (TMF_DBG(p) == 1) gmask = 0;
  const int id0 = p.item_offset + col0;
  if (gmask == 0) return;
  // Pass 2, common case: every lane's queue has room for all it can add (8 per hit group of its own row), so the survivors are
  // pushed straight from the registers pass 1 holds -- the loop is unrolled, each hit group is its own warp-uniform block with
  // static register indices, no call sites (nothing is live across a call) and no second accumulator read (each of those cost a
  // serialised ~300-cycle round trip while the MMA threads wait for the buffer).
  if (__all_sync(0xffffffffu, st.cq + 8 * __popc(hm) <= QCAP)) {
    int cq = st.cq;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (gmask & (1u << g)) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = (g & 1) * 8 + j;
          const float x = __uint_as_float(g < 2 ? ra[e] : g < 4 ? rb[e] : g < 6 ? rc[e] : rd[e]);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %0, %1;\n\t@p st.shared.v2.f32 [%2], {%0, %3};\n\t}"
                       ::"f"(x), "f"(thr), "r"(queue + 8u * (uint32_t)cq), "f"(__int_as_float(id0 + 8 * g + j)) : "memory");
          cq += (x >= thr) ? 1 : 0;
        }
      }
    }
    st.cq = cq;
    return;
  }
  // some lane may overflow its queue: re-read the hit groups from TMEM one at a time, draining the queues in between
  while (gmask) {  // warp-uniform
    const int g = __ffs(gmask) - 1;
    gmask &= gmask - 1;
    uint32_t v[8];
    tmem_ld8(t_base + (uint32_t)(8 * g), v);
    if (__any_sync(0xffffffffu, st.cq > QCAP - 8)) drain_queues(st, queue, buf, hrow, thr_slot, p.k, p.clamp);
    tmem_ld_wait_for8(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = __uint_as_float(v[j]);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %0, %1;\n\t@p st.shared.v2.f32 [%2], {%0, %3};\n\t}"
                   ::"f"(x), "f"(thr), "r"(queue + 8u * (uint32_t)st.cq), "f"(__int_as_float(id0 + 8 * g + j)) : "memory");
      st.cq += (x >= thr) ? 1 : 0;
    }
  }
}

// exact k-th largest approximate score of a row's list (n entries in global memory, streamed through L2) by an
// 8-bit-per-pass radix select; also returns the list maximum.  radix: 256 ints of shared-memory scratch.
__device__ __noinline__ float warp_select_kth(const float2* buf, int n, int k, int* radix, float& mx_out) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  float mx = -INFINITY;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) radix[lane * 8 + i] = 0;
    __syncwarp();
    for (int e0 = 0; e0 < n; e0 += 32 * 8) {  // 8 independent loads in flight per lane (one L2 round trip per 256 entries)
      float sc[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = e0 + 32 * t + lane;
        sc[t] = (e < n) ? __ldcg(&buf[e].x) : -INFINITY;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (e0 + 32 * t + lane < n) {
          mx = fmaxf(mx, sc[t]);
          const uint32_t key = f2key(sc[t]);
          if ((key & mask) == prefix) smem_inc(&radix[(key >> shift) & 255u]);
        }
      }
    }
    __syncwarp();
    int c[8];
    int lsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // lane L owns bins 255-8L .. 248-8L, visited in descending order
      c[i] = radix[255 - 8 * lane - i];
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  mx_out = mx;
  return key2f(prefix);
}

// Warp-cooperative compaction of one (row, group) list that is about to fill: entries below the row's current threshold
// (they can no longer be in the answer) are dropped in place, in ONE streaming pass (clamp mode also keeps the k lowest item
// ids: the zero-score fillers).  The histogram is not touched: dropped entries sit below the threshold bin, which the
// threshold scan never reaches again.  Returns the new length (all lanes).
__device__ __noinline__ int warp_compact_row(float2* buf, int n, float thr, int k, int clamp, int item_offset) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int base = 0;
  for (int b0 = 0; b0 < n; b0 += 32 * 8) {
    float2 x[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = b0 + 32 * t + lane;
      x[t] = (e < n) ? __ldcg(buf + e) : make_float2(-INFINITY, 0.f);
    }
    __syncwarp();  // every read of this batch precedes its writes (writes land at or below b0)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = b0 + 32 * t + lane;
      const bool keep = e < n && (x[t].x >= thr || (clamp && __float_as_int(x[t].y) - item_offset < k));
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcg(buf + base + __popc(bal & lt), x[t]);
      base += __popc(bal);
    }
    __syncwarp();
  }
  return base;
}

// ------------------------------------------------------------------ the fused kernel
// dynamic shared memory besides the B ring: alignment slack, A tile, 320 B of barriers + TMEM slot, per-virtual-row histograms,
// queues, published thresholds, rebuild outputs
static size_t topk_smem_fixed_bytes(int kb) {
  return 1024 + (size_t)kb * A_SUB_BYTES + 320 + (size_t)2 * BM * HSTRIDE * 4 + 16 + (size_t)2 * BM * 16 + (size_t)2 * BM * 8 +
         (size_t)NGRP * 4 * 4 + (size_t)NGRP * BM * QCAP * 8;
}

// CG2: the CTAs of a 2-CTA cluster work as a pair (tcgen05 cta_group::2): CTA r scores user block 2j + r, holds its own A tile
// and HALF of every B stage (64 of the tile's 128 items), and one M = 256 MMA issued by the leader fills both CTAs' TMEM.  A B
// stage is then 8 KB per SM for a full tile's worth of tensor work: with one CTA per SM the stage ring is what limits the
// pipe (measured: 4 x 16 KB stages in flight per SM and a ~1.5 us load round trip = 43 GB/s per SM, tensor pipe 33 % active).
template <bool DUMP, bool PROF, bool CG2>
__global__ void __launch_bounds__(TOPK_THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmapU, const __grid_constant__ CUtensorMap tmapV, const TopkParams p) {
  constexpr int B_STAGE = CG2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;  // bytes of a B stage in THIS CTA's shared memory
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;               // 0 = leader (issues the MMAs)
  const int cta_stride = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int cta_first = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  // user blocks of this CTA: first_ub, first_ub + ub_step, ...   (CG2: pair j takes blocks 2j and 2j + 1)
  const int first_ub = CG2 ? 2 * cta_first + (int)crank : cta_first;
  const int ub_step = CG2 ? 2 * cta_stride : cta_stride;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve (1024-byte aligned operand tiles first)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                   // kb sub-tiles of [128][64] 16-bit
  unsigned char* sB = sA + p.kb * A_SUB_BYTES;                // nstages x [BN (CG2: BN/2)][64] 16-bit
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.nstages * B_STAGE);
  uint64_t* full_bar = bars;                        // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;          // [MAX_STAGES]
  uint64_t* a_full = bars + 2 * MAX_STAGES;         // [1]
  uint64_t* a_empty = a_full + 1;                   // [1]
  uint64_t* tfull = a_empty + 1;                    // [NBUF]
  uint64_t* tempty = tfull + NBUF;                  // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + NBUF);
  uint32_t* hist_rows = reinterpret_cast<uint32_t*>(bars + 40);                  // [2][BM][HSTRIDE] per-row score histograms (by user-block parity)
  float4* rowp = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(hist_rows + 2 * BM * HSTRIDE) + 15) & ~(uintptr_t)15);  // [2][BM] (lo, w, 1/w, -)
  uint2* thr_sh = reinterpret_cast<uint2*>(rowp + 2 * BM);                       // [2][BM] (threshold key, user block it belongs to)
  int* done_sh = reinterpret_cast<int*>(thr_sh + 2 * BM);                        // [NGRP][4] last user block each epilogue warp has finished
  float2* queues = reinterpret_cast<float2*>(done_sh + NGRP * 4);                // [NGRP * BM][QCAP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto now = [] () -> long long { return PROF ? clock64() : 0ll; };  // cycle counters only in the profiling build

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapU) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapV) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nstages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(a_full), 1);
    mbar_init(smem_u32(a_empty), 2);  // both MMA-issuing threads commit it
    for (int s = 0; s < NBUF; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), (CG2 ? 2 : 1) * 4); }  // tempty: ONE arrival per epilogue warp (of both CTAs): 32 per-lane arrivals on one barrier word serialise (~300 cycles per warp and tile)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: all 512 columns = four 128x128 fp32 accumulators (one CTA per SM)
    if constexpr (CG2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = threadIdx.x; i < 2 * BM; i += TOPK_THREADS) thr_sh[i] = make_uint2(0u, 0xffffffffu);
  if (threadIdx.x < NGRP * 4) done_sh[threadIdx.x] = -1;
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The B stages form TWO rings of nstages / 2: ring r carries the tiles nt % 2 == r, the tiles of MMA thread r.  With one ring and
    // two consumers that each skip the other's stages, a consumer that is held up (its epilogue group is busy) could be lapped
    // twice by the ring and would then mistake a later phase of a stage's barrier for the one it waits for.
    if (lane == 0) {
      const int hs = p.nstages >> 1;
      int st_r[2] = {0, 0};
      uint32_t ph_r[2] = {0, 0};
      uint32_t a_phase = 0;
      for (int ub = first_ub; ub < p.n_ublocks; ub += ub_step) {
        long long t0 = now();
        mbar_wait_ctrl(smem_u32(a_empty), a_phase ^ 1);  // previous user block's MMAs retired
        long long w_empty = 0, w_aempty = now() - t0;
        if (!CG2 || crank == 0) mbar_expect_tx(smem_u32(a_full), (CG2 ? 2 : 1) * p.kb * A_SUB_BYTES);  // the pair's A tiles complete on the leader's barrier
        for (int kb = 0; kb < p.kb; ++kb) {
          if constexpr (CG2) tma_load_2d_cg2(smem_u32(sA + kb * A_SUB_BYTES), &tmapU, kb * BK, (p.ub0 + ub) * BM, smem_u32(a_full));
          else tma_load_2d(smem_u32(sA + kb * A_SUB_BYTES), &tmapU, kb * BK, (p.ub0 + ub) * BM, smem_u32(a_full));
        }
        a_phase ^= 1;
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          const int r = nt & 1;
          int rs = r ? st_r[1] : st_r[0];
          uint32_t phase = r ? ph_r[1] : ph_r[0];
          for (int kb = 0; kb < p.kb; ++kb) {
            const int stage = r * hs + rs;
            t0 = now();
            mbar_wait_ctrl(smem_u32(&empty_bar[stage]), phase ^ 1);
            w_empty += now() - t0;
            if (TMF_DBG(p) == 3) {  // diagnostic: no B loads at all (MMAs read whatever the stage holds)
              if (!CG2 || crank == 0) mbar_arrive(smem_u32(&full_bar[stage]));
            } else if constexpr (CG2) {  // this CTA's half of the tile (items nt * 128 + 64 * rank ...), bytes of both halves land on the leader's barrier
              if (crank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), B_STAGE_BYTES);
              tma_load_2d_cg2(smem_u32(sB + stage * B_STAGE), &tmapV, kb * BK, nt * BN + (int)crank * (BN / 2), smem_u32(&full_bar[stage]));
            } else {
              mbar_expect_tx(smem_u32(&full_bar[stage]), B_STAGE_BYTES);
              tma_load_2d(smem_u32(sB + stage * B_STAGE), &tmapV, kb * BK, nt * BN, smem_u32(&full_bar[stage]));
            }
#if TMF_PF_TILES > 0
            if (nt + TMF_PF_TILES < p.n_tiles)  // warm the L2 for a tile further down the sweep (no shared memory needed)
              asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmapV), "r"(kb * BK), "r"((nt + TMF_PF_TILES) * BN) : "memory");
#endif
            if (++rs == hs) { rs = 0; phase ^= 1; }
          }
          if (r) { st_r[1] = rs; ph_r[1] = phase; } else { st_r[0] = rs; ph_r[0] = phase; }
        }
        if (PROF) { atomicAdd(p.prof + 0, (unsigned long long)w_empty); atomicAdd(p.prof + 1, (unsigned long long)w_aempty); }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (CG2: of the leader CTA only) =====================
    // TWO issuing threads (warps 1 and 3, on different SM sub-partitions): thread r issues the tiles nt % 2 == r (accumulator
    // buffers r and r + 2) from its own stage ring.  One thread needs ~150 cycles per tcgen05.mma (descriptor arithmetic, the moves
    // into the uniform registers the instruction reads, the issue itself -- measured: a single issuer spent 77 % of its time
    // issuing with the tensor pipe 33 % active) against 64 cycles of execution, so one issuer starves the pipe.  Both threads
    // commit the end-of-block "A tile free" barrier (count 2).
    if (lane == 0 && crank == 0) {
      const int issuer = warp == 1 ? 0 : 1;
      const int hs = p.nstages >> 1;
      int rs = 0;  // position in this thread's stage ring
      uint32_t phase = 0, a_phase = 0;
      uint32_t acc_bits = 0;  // bit a = parity of the number of times accumulator a has been filled
      // operand format bits of the instruction descriptor: a_format (bit 7) / b_format (bit 10): 1 = bf16, 0 = fp16
      const bool f16 = p.force_fmt >= 0 ? p.force_fmt == 1 : use_fp16(p.fmt_stats);
      constexpr uint32_t kId = umma_idesc_bf16(CG2 ? 2 * BM : BM, BN);
      const uint32_t idesc = f16 ? (kId & ~((1u << 7) | (1u << 10))) : kId;
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));  // + (byte offset >> 4) selects a sub-tile / k-slice (no carry out of the 14-bit field)
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB));
      for (int ub = first_ub; ub < p.n_ublocks; ub += ub_step) {
        mbar_wait_ctrl(smem_u32(a_full), a_phase);
        a_phase ^= 1;
        long long w_tempty = 0, w_full = 0;
        const long long cb = now();
        for (int nt = issuer; nt < p.n_tiles; nt += 2) {
          const int acc = nt & (NBUF - 1);
          long long t0 = now();
          mbar_wait_ctrl(smem_u32(&tempty[acc]), ((acc_bits >> acc) & 1u) ^ 1u);  // the group drained this accumulator's previous tile
          w_tempty += now() - t0;
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < p.kb; ++kb) {
            const int stage = issuer * hs + rs;
            t0 = now();
            mbar_wait_ctrl(smem_u32(&full_bar[stage]), phase);
            w_full += now() - t0;
            tcgen05_fence_after();
            const uint64_t a_desc = a_desc0 + (uint64_t)((kb * A_SUB_BYTES) >> 4);
            const uint64_t b_desc = b_desc0 + (uint64_t)((stage * B_STAGE) >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (TMF_DBG(p) == 4 && (kb | k) != 0) continue;  // diagnostic: one MMA per tile (loads at full rate)
              if constexpr (CG2)
                tcgen05_mma_f16_cg2(d_tmem, a_desc + (uint64_t)(k * (UMMA_K * 2 / 16)), b_desc + (uint64_t)(k * (UMMA_K * 2 / 16)), idesc,
                                    (uint32_t)((kb | k) != 0));
              else
                tcgen05_mma_f16(d_tmem, a_desc + (uint64_t)(k * (UMMA_K * 2 / 16)), b_desc + (uint64_t)(k * (UMMA_K * 2 / 16)), idesc,
                                (uint32_t)((kb | k) != 0));
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if constexpr (CG2) tcgen05_commit_cg2(smem_u32(&empty_bar[stage])); else tcgen05_commit(smem_u32(&empty_bar[stage]));
            if (++rs == hs) { rs = 0; phase ^= 1; }
          }
          // accumulator ready for its epilogue group (of both CTAs of a pair)
          if constexpr (CG2) tcgen05_commit_cg2(smem_u32(&tfull[acc])); else tcgen05_commit(smem_u32(&tfull[acc]));
          acc_bits ^= 1u << acc;
        }
        if constexpr (CG2) tcgen05_commit_cg2(smem_u32(a_empty)); else tcgen05_commit(smem_u32(a_empty));  // this thread's MMAs no longer read the A tile
        if (PROF) {
          atomicAdd(p.prof + 2, (unsigned long long)w_tempty); atomicAdd(p.prof + 3, (unsigned long long)w_full);
          atomicAdd(p.prof + 21, (unsigned long long)(now() - cb)); atomicAdd(p.prof + 22, (unsigned long long)((p.n_tiles - issuer + 1) / 2));
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: threshold filter + survivor queues + SIMD list maintenance =====================
    // 16 warps = 4 groups of 4 (one warp per TMEM lane quarter).  Group g owns accumulator buffer g: it filters the tiles
    // nt % 4 == g, 128 columns in two 64-column passes.  A buffer is handed back when the 4 warps of its group (8 in a CTA pair) are
    // through with it and three other buffers are in flight meanwhile, so one slow warp (a queue drain, a list compaction) does not
    // stall the MMA threads.  (Measured alternative: two 256-column buffers with all 16 warps on every tile -- each hand-over then
    // waits for the slowest of 32 warps with one tile of slack, and the MMA threads waited for buffers half of the time.)
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;           // TMEM lane quarter == warp % 4
    const int trow = q * 32 + lane;   // row of the CTA's user tile
    const uint32_t queue = smem_u32(queues + (grp * BM + trow) * QCAP);
    const int n_items = (int)p.n_items;
    uint32_t acc_phase = 0;  // parity of the number of tiles this group has consumed
    int n_ub = 0;            // user blocks this CTA has started: parity selects the histogram / threshold buffers
    for (int ub = first_ub; ub < p.n_ublocks; ub += ub_step, ++n_ub) {
      const int par = n_ub & 1;
      const uint32_t hrow = smem_u32(hist_rows + (par * BM + trow) * HSTRIDE);
      const uint32_t thr_slot = smem_u32(thr_sh + par * BM + trow);  // .x = threshold key, .y = tag
      const long long lrow = (long long)ub * BM + trow;                // batch-local row (candidate buffers)
      const long long row = (long long)p.ub0 * BM + lrow;              // global row
      const bool valid = row < p.n_users;
      RowState st;
      st.thr = valid ? -INFINITY : INFINITY;
      st.lo = INFINITY; st.w = 0.f; st.inv_w = 0.f;
      st.E = p.erow[row];
      st.cnt = 0; st.cq = 0;
      // External bound (item-sharded scoring): B = a lower bound of the row's k-th best CANONICAL score over all
      // slabs, so a member of the global top-k has s~ >= B - E.  Clamp mode: only a positive bound says anything
      // (with B <= 0 the zero-score fillers matter).  Such a row filters at that floor from its first tile on.
      st.thr_ext = -INFINITY;
      if (!DUMP && p.row_bound != nullptr && valid) {
        const float B = p.row_bound[row];
        if (B > -INFINITY && (!p.clamp || B > 0.f)) st.thr_ext = B - st.E;
      }
      if (valid) st.thr = st.thr_ext;
      float2* buf = p.cand + (lrow * NGRP + grp) * CAPG;
      bool ready = false;  // this warp holds the row state of this user block (bin range, shared threshold)
      long long w_tfull = 0, w_work = 0, w_maint = 0;
      int n_own = 0;
      for (int nt = grp; nt < p.n_tiles; nt += NGRP, ++n_own) {
        const long long c0 = now();
        mbar_wait_epi(smem_u32(&tfull[grp]), acc_phase);
        const long long c1 = now();
        w_tfull += c1 - c0;
        acc_phase ^= 1u;
        __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the per-lane spin
        tcgen05_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * BN);
        const int col0 = nt * BN;
        if (DUMP) {
          uint32_t ra[32];
#pragma unroll 1
          for (int ch = 0; ch < BN / 32; ++ch) {
            tmem_ld32(t_base + (uint32_t)(ch * 32), ra);
            tmem_ld_wait_for(ra);
            dump_chunk(ra, col0 + ch * 32, n_items, valid, row, p);
          }
        } else if (nt == 0) {
          // ---- the row's FIRST tile (group 0): everything at or above the external floor is appended; mean and standard
          // deviation of its 128 scores give the histogram's bin range; a second read of the (still resident) accumulator bins
          // the appended scores.  The other groups wait for the published row before their first tile.
          // The buffers of this parity were last used two user blocks ago: every warp of this lane quarter must be past that.
          if (lane < NGRP)
            while ((int)lds_u32_volatile(smem_u32(done_sh + lane * 4 + q)) < n_ub - 2) __nanosleep(64);
          __syncwarp();
          float mx = -INFINITY, sum = 0.f, sumsq = 0.f;
          int nv = 0;
          uint32_t ra[32];
#pragma unroll 1
          for (int ch = 0; ch < BN / 32; ++ch) {
            tmem_ld32(t_base + (uint32_t)(ch * 32), ra);
            tmem_ld_wait_for(ra);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int col = col0 + ch * 32 + jj;
              const float x = __uint_as_float(ra[jj]);
              if (col < n_items) {
                mx = fmaxf(mx, x);
                sum += x;
                sumsq = fmaf(x, x, sumsq);
                ++nv;
                if (valid && x >= st.thr_ext) {
                  __stcg(buf + st.cnt, make_float2(x, __int_as_float(p.item_offset + col)));
                  ++st.cnt;
                }
              }
            }
          }
          {
            // bin range [mean, mean + W): W = 6.5 standard deviations of the sample (the k-th best of 10^6 .. 10^9 Gaussian scores
            // sits 3.7 .. 5.5 sigma above the mean), and never less than 2.5 x (sample maximum - mean) for heavy-tailed rows.  A
            // range from the sample maximum alone is fragile: one row in ~10^5 has a sample maximum below 1 sigma, its top bin
            // then saturates far below the k-th best score, its lists fill up and the row falls back to the exact path.
            const float mean = nv > 0 ? sum / (float)nv : 0.f;
            const float var = nv > 1 ? fmaxf(sumsq / (float)nv - mean * mean, 0.f) : 0.f;
            float W = fmaxf(6.5f * sqrtf(var), 2.5f * (mx - mean));
            if (!(W > 0.f) || !(W < 3e38f)) W = fmaxf(fabsf(mean), 1.0f) * 1e-3f;
            st.lo = mean; st.w = W / (float)NBINS; st.inv_w = (float)NBINS / W;
            for (int b2 = 0; b2 < HSTRIDE; ++b2) sts_u32(hrow + 4u * (uint32_t)b2, 0u);
            rowp[par * BM + trow] = make_float4(st.lo, st.w, st.inv_w, 0.f);
            __threadfence_block();
            thr_sh[par * BM + trow] = make_uint2(f2key(st.thr), (uint32_t)ub);  // publishes the row: the other groups may start
          }
          ready = true;
#pragma unroll 1
          for (int ch = 0; ch < BN / 32; ++ch) {
            tmem_ld32(t_base + (uint32_t)(ch * 32), ra);
            tmem_ld_wait_for(ra);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const float x = __uint_as_float(ra[jj]);
              if (col0 + ch * 32 + jj < n_items && valid && x >= st.thr_ext && x >= st.lo) {
                const int b2 = (int)fminf((x - st.lo) * st.inv_w, (float)(NBINS - 1));
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hrow + 4u * (uint32_t)b2) : "memory");
              }
            }
          }
          {
            const int bthr = hist_threshold_bin(hrow, p.k);  // warp-collective (votes): every lane calls it, `valid` only masks the use
            if (valid && bthr >= 0) {
              const float t = fmaxf(edge_threshold(st.lo, st.w, bthr, st.E, p.clamp), st.thr_ext);
              if (t > st.thr) {
                st.thr = t;
                asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(thr_slot), "r"(f2key(t)) : "memory");
              }
            }
          }
        } else {
          if (!ready) {  // first own tile of this user block: wait for group 0 to publish the row, adopt its bin range
            while (lds_u32_volatile(thr_slot + 4u) != (uint32_t)ub) __nanosleep(32);
            __syncwarp();
            __threadfence_block();
            const float4 rp = rowp[par * BM + trow];
            st.lo = rp.x; st.w = rp.y; st.inv_w = rp.z;
            ready = true;
          }
          if (st.thr < INFINITY) st.thr = fmaxf(st.thr, key2f(lds_u32_volatile(thr_slot)));  // the row's best threshold so far
          if (TMF_DBG(p) < 2 || TMF_DBG(p) == 5) {
            epilogue_tile(t_base, col0, st, queue, buf, hrow, thr_slot, p);
            epilogue_tile(t_base + (uint32_t)QN, col0 + QN, st, queue, buf, hrow, thr_slot, p);
          }
        }
        // the buffer is drained: hand it back to the MMA thread before any list maintenance (one arrival per warp)
        tcgen05_fence_before();
        __syncwarp();  // every lane's accumulator reads are complete
        if (lane == 0) {
          if constexpr (CG2) mbar_arrive_leader(smem_u32(&tempty[grp])); else mbar_arrive(smem_u32(&tempty[grp]));
        }
        const long long c2 = now();
        w_work += c2 - c1;
        if (!DUMP) {
          if (ready) {
            pop_one(st, queue, buf, hrow);
            // threshold refresh every 16th own tile, staggered over the warps
            if (((n_own + (warp - 4)) & 15) == 0) st.thr = refresh_threshold(st.thr, st.thr_ext, st.lo, st.w, st.E, hrow, thr_slot, p.k, p.clamp);
          }
          if (__any_sync(0xffffffffu, st.cq > QCAP / 2)) drain_queues(st, queue, buf, hrow, thr_slot, p.k, p.clamp);
          // a list about to fill: drop what the row's threshold has overtaken
          unsigned need = __ballot_sync(0xffffffffu, valid && st.cnt <= CAPG && st.cnt > CAPG - 4 * QCAP);
          if (need) {
            drain_queues(st, queue, buf, hrow, thr_slot, p.k, p.clamp);
            if (st.thr < INFINITY) st.thr = fmaxf(st.thr, key2f(lds_u32_volatile(thr_slot)));
            need = __ballot_sync(0xffffffffu, valid && st.cnt <= CAPG && st.cnt > CAPG - 4 * QCAP);
          }
          while (need) {
            const int owner = __ffs(need) - 1;
            need &= need - 1;
            const int n_o = __shfl_sync(0xffffffffu, st.cnt, owner);
            const float thr_o = __shfl_sync(0xffffffffu, st.thr, owner);
            float2* obuf = p.cand + (((long long)ub * BM + q * 32 + owner) * NGRP + grp) * CAPG;
            const int n_new = warp_compact_row(obuf, n_o, thr_o, p.k, p.clamp, p.item_offset);
            if (lane == owner) {
              st.cnt = n_new;
              if (st.cnt > CAPG - 8 * QCAP) {  // cannot shrink (massive ties): hand the row to the exact path
                st.cnt = CAPG + 1;
                st.thr = INFINITY;
              }
            }
            __syncwarp();
          }
        }
        w_maint += now() - c2;
      }
      if (PROF && lane == 0) {
        atomicAdd(p.prof + 4, (unsigned long long)w_tfull); atomicAdd(p.prof + 5, (unsigned long long)w_work);
        atomicAdd(p.prof + 6, (unsigned long long)w_maint); atomicAdd(p.prof + 7, (unsigned long long)n_own);
      }
      if (!DUMP) {
        drain_queues(st, queue, buf, hrow, thr_slot, p.k, p.clamp);
        const bool ovf = st.cnt > CAPG;
        float thr_fin = st.thr;
        if (ready && !ovf) thr_fin = fmaxf(thr_fin, key2f(lds_u32_volatile(thr_slot)));
        p.cnt[lrow * NGRP + grp] = valid ? (ovf ? -1 : st.cnt) : 0;
        p.thr_out[lrow * NGRP + grp] = ovf ? -INFINITY : thr_fin;
      }
      __syncwarp();
      if (lane == 0) { __threadfence_block(); done_sh[grp * 4 + q] = n_ub; }  // this warp no longer touches the buffers of this parity
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();  // the peer may still signal this CTA's barriers / read its shared memory until here
  if (warp == 2) {
    if constexpr (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ final selection + canonical rerank
constexpr int RR_WARPS = 8;
constexpr int SEL_CAP = 512;   // survivors per row the rerank can rank; more -> exact path
constexpr int RR_CHAINS = 4;   // independent fp64 FMA chains per lane

struct RerankParams {
  long long n_users, n_items, row0;  // rows [row0, row0 + n_rows) are batch-local rows [0, n_rows)
  long long n_rows;
  int k, clamp, item_offset, r, ld;
  const float* U;
  const float* V;
  const float* erow;
  const float2* cand;
  const int* cnt;
  const float* thr;
  int* ovf_count;
  int* ovf_rows;
  unsigned long long* prof;
  int* out_idx;
  float* out_score;
};

// k-th largest of m (<= SEL_CAP) scores held in shared memory pairs (score, id): 8-bit radix select, one warp
__device__ __forceinline__ float smem_select_kth(const float2* pr, int m, int k, int* radix) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0, mask = 0;
  int krem = k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) radix[lane * 8 + i] = 0;
    __syncwarp();
    for (int e = lane; e < m; e += 32) {
      const uint32_t key = f2key(pr[e].x);
      if ((key & mask) == prefix) smem_inc(&radix[(key >> shift) & 255u]);
    }
    __syncwarp();
    int c[8];
    int lsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // lane L owns bins 255-8L .. 248-8L, visited in descending order
      c[i] = radix[255 - 8 * lane - i];
      lsum += c[i];
    }
    int incl = lsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned reach = __ballot_sync(0xffffffffu, incl >= krem);
    const int F = reach ? __ffs(reach) - 1 : 31;
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

constexpr int STG_COLS = 64;              // fp32 components per staged slice of an item row
constexpr int STG_STRIDE = STG_COLS + 4;  // floats; 272-byte rows keep the per-lane float4 reads conflict-free

// One warp per user row.
//  1. the row's candidate list is streamed ONCE from global memory; entries that pass the main kernel's final
//     threshold (a superset of the answer, ~2k of them) land in shared memory;
//  2. the exact k-th largest approximate score of that superset tightens the threshold (k-th - 2E) -> ~1.5k candidates;
//  3. the candidates' fp32 item rows are staged through shared memory by per-lane bulk async copies (one 256-byte
//     copy per candidate and slice: 8 KB in flight per warp, where the per-lane strided loads of the first version
//     exposed a full DRAM latency every 16 bytes), and every lane runs ONE candidate's fp64 FMA chain in
//     component order -- bit-identical to oracle.canonical_scores;
//  4. (score, id) are packed into one 64-bit key and ranked by counting (branch-free).
__global__ void __launch_bounds__(RR_WARPS * 32) rerank_kernel(const RerankParams p) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = (size_t)p.ld * sizeof(double) + SEL_CAP * 8 + 32 * STG_STRIDE * 4 + 16;
  unsigned char* base = rr_smem + (size_t)w * per_warp;
  double* ud = reinterpret_cast<double*>(base);                       // [ld] this user's row, widened once
  float2* pr = reinterpret_cast<float2*>(ud + p.ld);                  // [SEL_CAP] (approx score, id), later 64-bit rank keys
  float* stg = reinterpret_cast<float*>(pr + SEL_CAP);                // [32][STG_STRIDE] staged item-row slices
  int* radix = reinterpret_cast<int*>(stg);                           // radix-select scratch before staging starts
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg + 32 * STG_STRIDE); // this warp's copy-completion barrier
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(pr);

  const long long lrow = (long long)blockIdx.x * RR_WARPS + w;
  if (lrow >= p.n_rows) return;
  const long long row = p.row0 + lrow;
  // the row's candidates: one list per epilogue group of the main kernel ([row][NGRP][CAPG]); each group's final threshold is
  // a valid keep-threshold for the whole row, so the highest one applies to all lists
  int ng[NGRP];
  int n = 0;
  bool overflowed = false;
  float thr = -INFINITY;
#pragma unroll
  for (int g = 0; g < NGRP; ++g) {
    ng[g] = p.cnt[lrow * NGRP + g];
    overflowed |= ng[g] < 0;
    n += max(ng[g], 0);
    thr = fmaxf(thr, p.thr[lrow * NGRP + g]);
  }
  if (overflowed) {  // a list saturated in the main kernel (massive ties): the exact path ranks this row
    if (lane == 0) {
      const int slot = atomicAdd(p.ovf_count, 1);
      p.ovf_rows[slot] = (int)row;
    }
    return;
  }
  float2* rowbuf = const_cast<float2*>(p.cand) + lrow * CAP;
  const int k = p.k;
  const unsigned lt = (1u << lane) - 1u;
  if (lane == 0) mbar_init(smem_u32(bar), 1);
  for (int c = lane; c < p.ld; c += 32) ud[c] = (double)p.U[row * p.ld + c];
  // ---- 1. superset by the main kernel's final threshold (or the clamp-mode filler rule)
  int m = 0;
#pragma unroll 1
  for (int g = 0; g < NGRP; ++g) {
    const float2* buf = rowbuf + g * CAPG;
    for (int b0 = 0; b0 < ng[g]; b0 += 256) {
      float2 x[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = b0 + 32 * t + lane;
        x[t] = (e < ng[g]) ? __ldcg(buf + e) : make_float2(-INFINITY, 0.f);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = b0 + 32 * t + lane;
        const bool keep = e < ng[g] && (x[t].x >= thr || (p.clamp && __float_as_int(x[t].y) - p.item_offset < k));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int pos = m + __popc(bal & lt);
        if (keep && pos < SEL_CAP) pr[pos] = x[t];
        m += __popc(bal);
      }
    }
  }
  __syncwarp();
  // ---- 2. exact k-th largest approximate score -> final keep-threshold, compact in place
  if (n > k && m > k) {  // (m <= k: an externally bounded row with few local candidates keeps them all)
    float kth;
    if (m <= SEL_CAP) {
      kth = smem_select_kth(pr, m, k, radix);
    } else {  // loose running thresholds (badly placed histograms): make the row's lists one contiguous list in place, select over it
      int off = ng[0];
#pragma unroll 1
      for (int g = 1; g < NGRP; ++g) {
        const float2* src = rowbuf + g * CAPG;
        for (int b0 = 0; b0 < ng[g]; b0 += 32) {  // destination <= source: forward copy, reads of a batch precede its writes
          const int e = b0 + lane;
          const float2 x = e < ng[g] ? __ldcg(src + e) : make_float2(0.f, 0.f);
          __syncwarp();
          if (e < ng[g]) __stcg(rowbuf + off + e, x);
        }
        off += ng[g];
        __syncwarp();
      }
      __threadfence_block();
      float mx;
      kth = warp_select_kth(rowbuf, n, k, radix, mx);
    }
    thr = fmaxf(thr, keep_threshold(kth, p.erow[row], p.clamp));
    if (m <= SEL_CAP) {
      int m2 = 0;
      for (int e0 = 0; e0 < m; e0 += 32) {
        const int e = e0 + lane;
        const float2 x = e < m ? pr[e] : make_float2(-INFINITY, 0.f);
        const bool keep = e < m && (x.x >= thr || (p.clamp && __float_as_int(x.y) - p.item_offset < k));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();  // all reads of this batch precede its writes (which land at or below e0)
        if (keep) pr[m2 + __popc(bal & lt)] = x;
        m2 += __popc(bal);
      }
      m = m2;
    } else {
      m = 0;
      for (int b0 = 0; b0 < n; b0 += 128) {
        float2 x[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = b0 + 32 * t + lane;
          x[t] = (e < n) ? __ldcg(rowbuf + e) : make_float2(-INFINITY, 0.f);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = b0 + 32 * t + lane;
          const bool keep = e < n && (x[t].x >= thr || (p.clamp && __float_as_int(x[t].y) - p.item_offset < k));
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          const int pos = m + __popc(bal & lt);
          if (keep && pos < SEL_CAP) pr[pos] = x[t];
          m += __popc(bal);
        }
      }
    }
  }
  if (m > SEL_CAP) {  // too many near-ties to rank here
    if (lane == 0) {
      if (p.prof) atomicAdd(p.prof + 9, 1ull);
      const int slot = atomicAdd(p.ovf_count, 1);
      p.ovf_rows[slot] = (int)row;
    }
    return;
  }
  __syncwarp();
  // ---- 3. canonical scores: 32 candidates per batch, item rows staged slice by slice
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  uint32_t parity = 0;
  const int n_slices = (p.ld + STG_COLS - 1) / STG_COLS;
  for (int t0 = 0; t0 < m; t0 += 32) {
    const int t = t0 + lane;
    const bool have = t < m;
    const int item = have ? __float_as_int(pr[t].y) : 0;
    const float* vrow = p.V + (long long)(item - p.item_offset) * p.ld;
    const int nb = min(32, m - t0);
    double acc = 0.0;
    for (int sl = 0; sl < n_slices; ++sl) {
      const int c0 = sl * STG_COLS;
      const int ncol = min(STG_COLS, p.ld - c0);  // multiple of 4
      // the previous slice's shared-memory reads (generic proxy) must be ordered before the async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_expect_tx(smem_u32(bar), (uint32_t)(nb * ncol * 4));
      __syncwarp();
      if (have) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stg + lane * STG_STRIDE)), "l"(vrow + c0), "r"(ncol * 4), "r"(smem_u32(bar))
                     : "memory");
      }
      mbar_wait(smem_u32(bar), parity);
      parity ^= 1;
      const float4* mine = reinterpret_cast<const float4*>(stg + lane * STG_STRIDE);
      const double2* u2 = reinterpret_cast<const double2*>(ud + c0);
#pragma unroll 4
      for (int c4 = 0; c4 < ncol / 4; ++c4) {
        const float4 x = mine[c4];
        const double2 ua = u2[2 * c4], ub = u2[2 * c4 + 1];
        acc = fma(ua.x, (double)x.x, acc);
        acc = fma(ua.y, (double)x.y, acc);
        acc = fma(ub.x, (double)x.z, acc);
        acc = fma(ub.y, (double)x.w, acc);
      }
    }
    if (have) {
      float sc = (float)acc;
      if (p.clamp) sc = sc > 0.f ? sc : 0.f;  // tf.where(p > 0, p, 0.0)
      sc = sc + 0.0f;                         // -0 -> +0 like the oracle
      keys[t] = ((unsigned long long)f2key(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)item);
    }
  }
  __syncwarp();
  // ---- 4. rank by counting: larger key = (higher score, then lower item id); ids are distinct so ranks are too
  for (int g0 = 0; g0 < m; g0 += 128) {
    unsigned long long mk[4];
    int rank[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = g0 + 32 * i + lane;
      mk[i] = t < m ? keys[t] : ~0ull;
      rank[i] = 0;
    }
    for (int o = 0; o < m; ++o) {
      const unsigned long long ko = keys[o];
#pragma unroll
      for (int i = 0; i < 4; ++i) rank[i] += ko > mk[i] ? 1 : 0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = g0 + 32 * i + lane;
      if (t < m && rank[i] < k) {
        p.out_idx[row * k + rank[i]] = (int)(0xffffffffu - (uint32_t)(mk[i] & 0xffffffffull));
        p.out_score[row * k + rank[i]] = key2f((uint32_t)(mk[i] >> 32));
      }
    }
  }
  // fewer than k local candidates (only with an external bound): pad with entries that lose every comparison
  for (int t = m + lane; t < k; t += 32) {
    p.out_idx[row * k + t] = 0x7fffffff;
    p.out_score[row * k + t] = -INFINITY;
  }
}

// exact path for rows whose candidate list overflowed (huge tie groups, adversarially ordered scores): all
// canonical scores of the row go to scratch; a block-wide radix select finds the k-th largest score T; entries > T
// plus the lowest-indexed entries == T are collected (k in total) and ranked by (score desc, id asc).
__global__ void __launch_bounds__(256) exact_rows_kernel(const RerankParams p, const int* __restrict__ ovf_count,
                                                         const int* __restrict__ ovf_rows, float* __restrict__ scratch) {
  __shared__ int hist[256];
  __shared__ int s_bin, s_krem, s_cnt_gt, s_cnt_eq, s_warp_tot[8];
  __shared__ float sel_s[128];
  __shared__ int sel_i[128];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_ovf = *ovf_count;
  const int n = (int)p.n_items;
  const int k = p.k;
  float* sc = scratch + (long long)blockIdx.x * p.n_items;
  for (int o = blockIdx.x; o < n_ovf; o += gridDim.x) {
    const long long row = ovf_rows[o];
    const float* u = p.U + row * p.ld;
    for (int i = tid; i < n; i += 256) {
      const float* v = p.V + (long long)i * p.ld;
      double acc = 0.0;
      for (int c = 0; c < p.r; ++c) acc = fma((double)u[c], (double)v[c], acc);
      float s = (float)acc;
      if (p.clamp) s = s > 0.f ? s : 0.f;
      sc[i] = s + 0.0f;
    }
    __syncthreads();
    // ---- k-th largest key, 8 bits per pass
    uint32_t prefix = 0, mask = 0;
    int krem = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += 256) {
        const uint32_t key = f2key(sc[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= krem) break;
          cum += hist[b];
        }
        s_bin = b;
        s_krem = krem - cum;
      }
      __syncthreads();
      prefix |= (uint32_t)s_bin << shift;
      mask |= 255u << shift;
      krem = s_krem;
      __syncthreads();
    }
    const float T = key2f(prefix);  // exactly the k-th largest score; krem of the entries == T are needed
    if (tid == 0) { s_cnt_gt = 0; s_cnt_eq = 0; }
    __syncthreads();
    // ---- collect: every entry > T (k - krem of them), and the krem lowest-indexed entries == T (ordered scan)
    for (int i0 = 0; i0 < n; i0 += 256) {
      const int i = i0 + tid;
      const float s = i < n ? sc[i] : -INFINITY;
      const bool gt = i < n && s > T;
      const bool eq = i < n && s == T;
      if (gt) {
        const int pos = atomicAdd(&s_cnt_gt, 1);
        sel_s[pos] = s;
        sel_i[pos] = i;
      }
      const unsigned beq = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp_tot[wid] = __popc(beq);
      __syncthreads();
      int before = s_cnt_eq;
      for (int w2 = 0; w2 < wid; ++w2) before += s_warp_tot[w2];
      const int my = before + __popc(beq & ((1u << lane) - 1u));
      if (eq && my < krem) {  // slots [k - krem, k) hold the tied entries in index order
        sel_s[k - krem + my] = s;
        sel_i[k - krem + my] = i;
      }
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w2 = 0; w2 < 8; ++w2) tot += s_warp_tot[w2];
        s_cnt_eq += tot;
      }
      __syncthreads();
    }
    // ---- rank the k selected entries
    if (tid < k) {
      const float s = sel_s[tid];
      const int id = sel_i[tid];
      int rank = 0;
      for (int j = 0; j < k; ++j) rank += (sel_s[j] > s) || (sel_s[j] == s && sel_i[j] < id);
      p.out_idx[row * k + rank] = id + p.item_offset;
      p.out_score[row * k + rank] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ operand format
// The tensor-core operands are 16-bit: bf16 (8 significant bits, fp32's exponent range) or fp16 (11 significant bits,
// normal range 6e-5 .. 65504).  Where the embeddings fit fp16's range its rounding residuals -- and with them the
// data-derived error bound E, the survivor counts of the epilogue and the rerank's candidate lists -- are ~8x smaller.
// operand_stats_kernel measures both roundings on the actual data; use_fp16() picks the format with the smaller
// relative residual (bf16 whenever a component would overflow fp16).  stats: [0] sum v^2, [1] sum (v - fp16(v))^2,
// [2] sum (v - bf16(v))^2, [3] max|v| (as int bits) for U; [4..7] the same for V.  Either choice is exact (E follows the
// chosen rounding); the float atomics only make a borderline CHOICE run-dependent, never a result.
__device__ __forceinline__ bool use_fp16(const float* stats) {
  if (stats == nullptr) return false;
  const float mx = fmaxf(__int_as_float(__float_as_int(stats[3])), __int_as_float(__float_as_int(stats[7])));
  if (!(mx < 60000.f)) return false;
  const float su = fmaxf(stats[0], 1e-30f), sv = fmaxf(stats[4], 1e-30f);
  const float e16 = sqrtf(stats[1] / su) + sqrtf(stats[5] / sv);
  const float ebf = sqrtf(stats[2] / su) + sqrtf(stats[6] / sv);
  return e16 < ebf;
}

__global__ void __launch_bounds__(256) operand_stats_kernel(const float* __restrict__ src, long long n, int r, int ld, float* __restrict__ stats) {
  float ss = 0.f, s16 = 0.f, sbf = 0.f, mx = 0.f;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {  // one warp per row
    const float4* p4 = reinterpret_cast<const float4*>(src + row * ld);
    for (int c4 = lane; c4 < (ld >> 2); c4 += 32) {  // pad columns [r, ld) are zero: they add nothing
      const float4 q = __ldg(p4 + c4);
      const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = vv[j];
        const float d16 = v - __half2float(__float2half_rn(v));
        const float dbf = v - __bfloat162float(__float2bfloat16_rn(v));
        ss = fmaf(v, v, ss);
        s16 = fmaf(d16, d16, s16);
        sbf = fmaf(dbf, dbf, sbf);
        mx = fmaxf(mx, fabsf(v));
      }
    }
  }
  (void)r;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    s16 += __shfl_xor_sync(0xffffffffu, s16, o);
    sbf += __shfl_xor_sync(0xffffffffu, sbf, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats + 0, ss);
    atomicAdd(stats + 1, s16);
    atomicAdd(stats + 2, sbf);
    atomicMax(reinterpret_cast<int*>(stats) + 3, __float_as_int(mx));  // non-negative floats order like ints (inf/NaN sort high)
  }
}

// ------------------------------------------------------------------ operand packing
// one warp per row: fp32 [n, ld] -> bf16 [n_pad, k_pad] (zero padded); per row the l2 norms of the row, of its bf16
// image and of the rounding residual; optional global maxima of the row norm and the residual norm
__global__ void pack_bf16_kernel(const float* __restrict__ src, long long n, int r, int ld, __nv_bfloat16* __restrict__ dst,
                                 long long n_pad, int k_pad, float* __restrict__ norms, float* __restrict__ norms_bf,
                                 float* __restrict__ norms_res, int* __restrict__ max_bits, int nan_pad,
                                 const float* __restrict__ fmt_stats, int force_fmt) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_pad) return;
  const bool f16 = force_fmt >= 0 ? force_fmt == 1 : use_fp16(fmt_stats);
  unsigned short* dst16 = reinterpret_cast<unsigned short*>(dst);
  float ss = 0.f, sb = 0.f, sd = 0.f;
  for (int c = lane; c < k_pad; c += 32) {
    float v = 0.f;
    if (row < n && c < r) v = src[row * ld + c];
    unsigned short bits;
    float vb;
    if (f16) {
      const __half h = __float2half_rn(v);
      bits = __half_as_ushort(h);
      vb = __half2float(h);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      bits = __bfloat16_as_ushort(h);
      vb = __bfloat162float(h);
    }
    const float d = v - vb;  // exact
    ss = fmaf(v, v, ss);
    sb = fmaf(vb, vb, sb);
    sd = fmaf(d, d, sd);
    // padded ITEM rows are NaN: their scores are NaN in every accumulator row, which fmaxf ignores and every
    // >= test rejects -- the epilogue needs no per-column bounds checks
    dst16[row * k_pad + c] = (nan_pad && row >= n) ? (unsigned short)(f16 ? 0x7E00 : 0x7FC0) : bits;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if (lane == 0) {
    const float nrm = sqrtf(ss) * NORM_SLACK, nb = sqrtf(sb) * NORM_SLACK, nd = sqrtf(sd) * NORM_SLACK;
    if (norms) norms[row] = nrm;
    if (norms_bf) norms_bf[row] = nb;
    if (norms_res) norms_res[row] = nd;
    if (max_bits) {  // non-negative floats order like ints
      atomicMax(max_bits, __float_as_int(nrm));
      atomicMax(max_bits + 1, __float_as_int(nd));
    }
  }
}

// E[row] = |du| max|v| + |u~| max|dv| + accumulation slack (see ACC_UNIT)
__global__ void row_error_kernel(const float* __restrict__ unorm_bf, const float* __restrict__ unorm_res, const int* __restrict__ vmax_bits,
                                 int k_pad, float* __restrict__ erow, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float vmax = __int_as_float(vmax_bits[0]), dvmax = __int_as_float(vmax_bits[1]);
  const float e = unorm_res[i] * vmax + unorm_bf[i] * dvmax + ACC_UNIT * (float)k_pad * unorm_bf[i] * vmax;
  erow[i] = e * NORM_SLACK + 1e-30f;
}

// ------------------------------------------------------------------ host side
static int make_tmap(CUtensorMap* map, void* base, long long rows, int k_pad, int box_rows) {
  return make_tmap_k64(map, base, rows, k_pad, box_rows);
}

struct TopkLayout {
  long long nu_pad, ni_pad, batch_rows;
  int k_pad, kb;
  size_t off_ub, off_vb, off_unorm, off_ures, off_erow, off_vnorm, off_vmax, off_cand, off_cnt, off_thr, off_ovfc, off_ovfr, off_scratch, total;
  int scratch_rows;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static TopkLayout topk_layout(long long n_users, long long n_items, int r) {
  TopkLayout L{};
  L.nu_pad = cdiv(n_users, 2 * BM) * (2 * BM);  // whole PAIRS of user blocks (the CTA pair of cta_group::2 takes two at a time)
  L.ni_pad = cdiv(n_items, BN) * BN;
  L.batch_rows = std::min<long long>(L.nu_pad, (long long)UB_BATCH * BM);
  L.k_pad = (int)(cdiv(r, BK) * BK);
  L.kb = L.k_pad / BK;
  size_t o = 0;
  L.off_ub = o; o = align_up(o + (size_t)L.nu_pad * L.k_pad * 2, 1024);
  L.off_vb = o; o = align_up(o + (size_t)L.ni_pad * L.k_pad * 2, 1024);
  L.off_unorm = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.off_ures = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.off_erow = o; o = align_up(o + (size_t)L.nu_pad * 4, 256);
  L.off_vnorm = o; o = align_up(o + (size_t)L.ni_pad * 4, 256);
  L.off_vmax = o; o += 256;
  L.off_cand = o; o = align_up(o + (size_t)L.batch_rows * CAP * 8, 256);
  L.off_cnt = o; o = align_up(o + (size_t)L.batch_rows * NGRP * 4, 256);
  L.off_thr = o; o = align_up(o + (size_t)L.batch_rows * NGRP * 4, 256);
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
                                                                                n_pad, k_pad, norms, nullptr, nullptr, nullptr, 0, nullptr, 0);
  TMF_LAUNCH_CHECK();
  return TMF_OK;
}

extern "C" size_t tmf_score_topk_ws_bytes(int64_t n_users, int64_t n_items, int32_t n_comp, int32_t k) {
  (void)k;
  return topk_layout(n_users, n_items, n_comp).total;
}

static int score_topk_impl(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                           int32_t k, int32_t clamp, int32_t item_offset, int32_t* out_idx, float* out_score, void* ws,
                           size_t ws_bytes, tmf_stream_t stream, float* dump, const float* row_bound = nullptr, int force_fmt = -1) {
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
  float* unorm = reinterpret_cast<float*>(w + L.off_unorm);  // |bf16(u)|
  float* ures = reinterpret_cast<float*>(w + L.off_ures);    // |u - bf16(u)|
  float* erow = reinterpret_cast<float*>(w + L.off_erow);
  float* vnorm = reinterpret_cast<float*>(w + L.off_vnorm);
  int* vmax_bits = reinterpret_cast<int*>(w + L.off_vmax);
  float2* cand = reinterpret_cast<float2*>(w + L.off_cand);
  int* cnt = reinterpret_cast<int*>(w + L.off_cnt);
  float* thr = reinterpret_cast<float*>(w + L.off_thr);
  int* ovfc = reinterpret_cast<int*>(w + L.off_ovfc);
  int* ovfr = reinterpret_cast<int*>(w + L.off_ovfr);
  float* scratch = reinterpret_cast<float*>(w + L.off_scratch);
  cudaStream_t st = as_stream(stream);

  TMF_CUDA(cudaMemsetAsync(vmax_bits, 0, 8, st));
  TMF_CUDA(cudaMemsetAsync(ovfc, 0, 4, st));
#ifdef TMF_DEVTOOLS
  if (force_fmt < 0) {
    const char* e = getenv("TMF_TOPK_FMT");  // A/B aid: "bf16" / "f16"
    if (e) force_fmt = (e[0] == 'f' || e[0] == 'h') ? 1 : 0;
  }
#endif
  float* fmt_stats = reinterpret_cast<float*>(w + L.off_vmax + 16);  // 8 floats inside the 256-byte scalar block
  if (force_fmt < 0) {
    TMF_CUDA(cudaMemsetAsync(fmt_stats, 0, 32, st));
    const int sgrid = 8 * kNumSMs;
    operand_stats_kernel<<<sgrid, 256, 0, st>>>(U, n_users, n_comp, ld, fmt_stats);
    operand_stats_kernel<<<sgrid, 256, 0, st>>>(V, n_items, n_comp, ld, fmt_stats + 4);
  }
  pack_bf16_kernel<<<(unsigned)cdiv(L.nu_pad * 32, 256), 256, 0, st>>>(U, n_users, n_comp, ld, Ub, L.nu_pad, L.k_pad, nullptr, unorm, ures,
                                                                       nullptr, 0, fmt_stats, force_fmt);
  pack_bf16_kernel<<<(unsigned)cdiv(L.ni_pad * 32, 256), 256, 0, st>>>(V, n_items, n_comp, ld, Vb, L.ni_pad, L.k_pad, vnorm, nullptr, nullptr,
                                                                       vmax_bits, 1, fmt_stats, force_fmt);
  row_error_kernel<<<(unsigned)cdiv(L.nu_pad, 256), 256, 0, st>>>(unorm, ures, vmax_bits, L.k_pad, erow, L.nu_pad);
  TMF_LAUNCH_CHECK();

  CUtensorMap tmapU, tmapV;
  int rc = make_tmap(&tmapU, Ub, L.nu_pad, L.k_pad, BM);
  if (rc) return rc;
  // CTA pairs (cta_group::2) unless switched off for an A/B (development builds only)
  bool cg2 = true;
#ifdef TMF_DEVTOOLS
  { const char* e = getenv("TMF_TOPK_CG2"); if (e) cg2 = atoi(e) != 0; }
#endif
  rc = make_tmap(&tmapV, Vb, L.ni_pad, L.k_pad, cg2 ? BN / 2 : BN);  // a pair's CTA loads its half of every item tile
  if (rc) return rc;

  TopkParams p{};
  p.n_users = n_users; p.n_items = n_items;
  p.n_tiles = (int)(L.ni_pad / BN); p.kb = L.kb;
  p.k = k; p.clamp = clamp ? 1 : 0; p.item_offset = item_offset;
  p.erow = erow;
  p.cand = cand; p.cnt = cnt; p.thr_out = thr;
  p.dump = dump; p.dump_ld = n_items;
  p.row_bound = row_bound;
  p.fmt_stats = fmt_stats; p.force_fmt = force_fmt;
  p.dbg = 0;
  p.prof = nullptr;
#ifdef TMF_DEVTOOLS
  { const char* e = getenv("TMF_TOPK_DEBUG"); p.dbg = e ? atoi(e) : 0; }
  if (getenv("TMF_TOPK_PROF")) {
    p.prof = reinterpret_cast<unsigned long long*>(w + L.off_vmax + 64);
    TMF_CUDA(cudaMemsetAsync(p.prof, 0, 192, st));
  }
#endif

  RerankParams q{};
  q.n_users = n_users; q.n_items = n_items; q.k = k; q.clamp = p.clamp; q.item_offset = item_offset; q.r = n_comp; q.ld = ld;
  q.U = U; q.V = V; q.erow = erow; q.prof = p.prof; q.cand = cand; q.cnt = cnt; q.thr = thr; q.ovf_count = ovfc; q.ovf_rows = ovfr;
  q.out_idx = out_idx; q.out_score = out_score;

  const size_t smem_max = 227 * 1024;
  const size_t fixed = topk_smem_fixed_bytes(L.kb);
  const size_t stage_bytes = cg2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
  TMF_REQUIRE(fixed + 2 * stage_bytes <= smem_max, "tmf_score_topk: shared memory exhausted (n_components too large)");
  p.nstages = (int)std::min<size_t>(MAX_STAGES, (smem_max - fixed) / stage_bytes);
  p.nstages &= ~1;  // two rings of nstages / 2 (even / odd item tiles)
  const size_t smem = fixed + (size_t)p.nstages * stage_bytes;
  auto launch_main = [&](int grid) -> int {
    void (*kern)(const CUtensorMap, const CUtensorMap, const TopkParams);
    if (dump != nullptr) kern = cg2 ? score_topk_kernel<true, false, true> : score_topk_kernel<true, false, false>;
    else if (p.prof) kern = cg2 ? score_topk_kernel<false, true, true> : score_topk_kernel<false, true, false>;
    else kern = cg2 ? score_topk_kernel<false, false, true> : score_topk_kernel<false, false, false>;
    TMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TOPK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TMF_CUDA(cudaLaunchKernelEx(&cfg, kern, tmapU, tmapV, p));
    return TMF_OK;
  };
  const size_t rr_smem = (size_t)RR_WARPS * ((size_t)ld * sizeof(double) + SEL_CAP * 8 + 32 * STG_STRIDE * 4 + 16);
  TMF_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));

  // users go through in batches of UB_BATCH blocks so the candidate workspace stays bounded; V stays packed
  const int total_ublocks = (int)(L.nu_pad / BM);
  for (int ub0 = 0; ub0 < total_ublocks; ub0 += UB_BATCH) {
    p.ub0 = ub0;
    p.n_ublocks = std::min(UB_BATCH, total_ublocks - ub0);
    // one persistent CTA per SM (all of its TMEM, ~220 KB of its shared memory); n_ublocks is even, CTA pairs need an even grid
    const int grid = std::min(kNumSMs, p.n_ublocks) & ~1;
    rc = launch_main(grid);
    if (rc) return rc;
    TMF_LAUNCH_CHECK();
    if (dump != nullptr) continue;
    q.row0 = (long long)ub0 * BM;
    q.n_rows = std::min<long long>((long long)p.n_ublocks * BM, n_users - q.row0);
    rerank_kernel<<<(unsigned)cdiv(q.n_rows, RR_WARPS), RR_WARPS * 32, rr_smem, st>>>(q);
    TMF_LAUNCH_CHECK();
  }
  if (dump != nullptr) return TMF_OK;
  exact_rows_kernel<<<L.scratch_rows, 256, 0, st>>>(q, ovfc, ovfr, scratch);
  TMF_LAUNCH_CHECK();
  if (p.prof) {  // profiling aid only: synchronises
    unsigned long long h[24]; int novf = 0;
    TMF_CUDA(cudaStreamSynchronize(st));
    TMF_CUDA(cudaMemcpy(h, p.prof, 192, cudaMemcpyDeviceToHost));
    TMF_CUDA(cudaMemcpy(&novf, ovfc, 4, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[tmf prof] per tile: MMA thread total %.0f, wait tempty %.0f, wait full %.0f cycles\n", (double)h[21] / h[22], (double)h[2] / h[22], (double)h[3] / h[22]);
    fprintf(stderr, "[tmf prof] producer wait empty %.3g, a_empty %.3g | mma wait tempty %.3g, full %.3g | epilogue (per warp-tile, n=%llu) "
                    "wait tfull %.3g, work %.3g, maintenance %.3g cycles | overflow rows %d (main %llu, rerank %llu) | tile-end drains: %llu, %.0f cycles each\n",
            (double)h[0], (double)h[1], (double)h[2], (double)h[3], h[7], (double)h[4] / h[7], (double)h[5] / h[7], (double)h[6] / h[7], novf, h[8], h[9], h[11], h[11] ? (double)h[10] / h[11] : 0.0);
    fprintf(stderr, "[tmf prof] rebuilds: first %llu x %.0f cycles, saturated %llu x %.0f, generic %llu x %.0f | pre-rebuild drains %llu x %.0f | appended entries %.4g (%.1f per row-sweep)\n",
            h[12], h[12] ? (double)h[13] / h[12] : 0.0, h[14], h[14] ? (double)h[15] / h[14] : 0.0, h[16], h[16] ? (double)h[17] / h[16] : 0.0,
            h[18], h[18] ? (double)h[19] / h[18] : 0.0, (double)h[20], h[7] ? (double)h[20] / (32.0 * h[7]) : 0.0);
  }
  return TMF_OK;
}

extern "C" int tmf_score_topk(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                              int32_t k, int32_t clamp, int32_t item_offset, int32_t* out_idx, float* out_score, void* ws,
                              size_t ws_bytes, tmf_stream_t stream) {
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, k, clamp, item_offset, out_idx, out_score, ws, ws_bytes, stream, nullptr);
}

extern "C" int tmf_score_topk_bounded(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                                      int32_t k, int32_t clamp, int32_t item_offset, const float* row_bound, int32_t* out_idx,
                                      float* out_score, void* ws, size_t ws_bytes, tmf_stream_t stream) {
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, k, clamp, item_offset, out_idx, out_score, ws, ws_bytes, stream, nullptr,
                         row_bound);
}

extern "C" int tmf_score_dense_bf16(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                                    float* P, void* ws, size_t ws_bytes, tmf_stream_t stream) {
  TMF_REQUIRE(P != nullptr, "tmf_score_dense_bf16: null output");
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, 1, 0, 0, nullptr, nullptr, ws, ws_bytes, stream, P, nullptr, 0);
}

extern "C" int tmf_score_dense_tc(const float* U, int64_t n_users, const float* V, int64_t n_items, int32_t n_comp, int32_t ld,
                                  int32_t operand_format, float* P, void* ws, size_t ws_bytes, tmf_stream_t stream) {
  TMF_REQUIRE(P != nullptr, "tmf_score_dense_tc: null output");
  TMF_REQUIRE(operand_format >= -1 && operand_format <= 1, "tmf_score_dense_tc: operand_format is -1 (auto), 0 (bf16) or 1 (fp16)");
  return score_topk_impl(U, n_users, V, n_items, n_comp, ld, 1, 0, 0, nullptr, nullptr, ws, ws_bytes, stream, P, nullptr, operand_format);
}
