// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), for the dense contractions of the embedding
// towers: the ReLU embedding's second stage H.W and its backward H^T.dE / dE.W^T (embedding_graphs.py:85-87) and X.W / X^T.dE
// when the feature matrix is genuinely dense (embedding_graphs.py:38).
//
// north_star's tolerance is 1e-5 relative in fp32, which one bf16 product (2^-8) cannot meet, so every fp32 operand is split
// EXACTLY into three bf16 planes  x = x1 + x2 + x3  (x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2): 3 x 8 mantissa
// bits cover fp32's 24) and the six products whose weight is >= 2^-16 relative are accumulated in fp32 in TMEM:
//     A.B^T  ~  A1.B1 + A1.B2 + A2.B1 + A1.B3 + A3.B1 + A2.B2          (dropped terms <= 3 * 2^-24 |a||b|)
// i.e. fp32-level accuracy at 1/6 of the bf16 tensor rate -- still several times the SIMT fp32 rate of gemm_f32_kernel.
//
// The tensor core's own fp32 accumulation is NOT round-to-nearest: measured on B200, a chain of n accumulating MMAs leaves an
// error that grows ~linearly with n (4-6e-7 sum|a||b| after the 24 MMAs of one 64-wide k-block, 2.6e-6 after 384, 5e-6 after
// 768) where IEEE fp32 adds would leave ~1e-7.  So a chain never runs longer than ONE k-block: every k-block is computed
// into a fresh TMEM accumulator (two of them, double-buffered) and the epilogue warps add it to fp32 register accumulators
// with ordinary round-to-nearest adds ("promotion", as FP8 GEMMs do) -- the error then stays at the one-block level however
// large K is.
//
//   split3_pack_kernel : fp32 [R, C] (row-major, optional transpose) -> three K-major bf16 planes [3][rows_pad][k_pad]
//   gemm_tc_kernel     : one 128 x BN output tile per CTA and K split; warp 0 = TMA producer (3 A + 3 B sub-tiles per
//                        64-wide k-block, 128-byte swizzle), warp 1 = single-thread tcgen05.mma issuer (24 MMAs per k-block),
//                        warp 2 = TMEM allocator, warps 4-7 = promotion + epilogue (tcgen05.ld of each k-block's tile -> fp32
//                        register accumulators, row per thread -> global)
//   splitk_reduce_kernel: fixed-order sum of the K-split partial tiles (deterministic)
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace tmf {

constexpr int GT_BM = 128;
constexpr int GT_BK = 64;
constexpr int GT_THREADS = 256;
constexpr int GT_A_SUB = GT_BM * GT_BK * 2;  // 16 KB per plane and k-block

// one warp per 32 x 32 tile: coalesced reads of the fp32 source, coalesced 16-bit writes of the three planes, either orientation
__global__ void __launch_bounds__(256) split3_pack_kernel(const float* __restrict__ src, long long R, long long C, long long ld, int transpose,
                                                          __nv_bfloat16* __restrict__ dst, long long rows_pad, long long k_pad) {
  __shared__ float tile[8][32][33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles_k = k_pad / 32, tiles_r = rows_pad / 32;
  const long long t = (long long)blockIdx.x * 8 + w;
  if (t >= tiles_k * tiles_r) return;
  const long long tr = t / tiles_k, tk = t % tiles_k;  // tile of the OUTPUT planes: rows [32 tr, +32), k [32 tk, +32)
  // output element (row, k) = src[row][k] (no transpose) or src[k][row] (transpose)
  float (*tl)[33] = tile[w];
  if (!transpose) {
    for (int i = 0; i < 32; ++i) {
      const long long row = tr * 32 + i, k = tk * 32 + lane;
      tl[i][lane] = (row < R && k < C) ? src[row * ld + k] : 0.f;
    }
  } else {
    for (int i = 0; i < 32; ++i) {  // read along the source's contiguous dimension (= output rows), transpose through shared memory
      const long long k = tk * 32 + i, row = tr * 32 + lane;
      tl[lane][i] = (k < R && row < C) ? src[k * ld + row] : 0.f;
    }
  }
  __syncwarp();
  const long long plane = rows_pad * k_pad;
  for (int i = 0; i < 32; ++i) {
    const float x = tl[i][lane];
    const __nv_bfloat16 b1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b1);        // exact
    const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b2);       // exact
    const __nv_bfloat16 b3 = __float2bfloat16_rn(r2);
    const long long o = (tr * 32 + i) * k_pad + tk * 32 + lane;
    dst[o] = b1;
    dst[plane + o] = b2;
    dst[2 * plane + o] = b3;
  }
}

struct GemmTcParams {
  long long M, N;        // real output size
  long long m_pad, n_pad;
  int kb_total;          // k-blocks of 64
  int splits;            // K splits (grid.z); split s covers k-blocks [s * kb_per, min((s+1) * kb_per, kb_total))
  int kb_per;
  int nstages;
  float* C;              // splits == 1: the output [M, ldc]; else the partial buffer [splits][M][N]
  long long ldc;
};

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmTcParams p) {
  constexpr int B_SUB = BN * GT_BK * 2;
  constexpr int STAGE = 3 * GT_A_SUB + 3 * B_SUB;
  extern __shared__ __align__(1024) unsigned char gsm_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * STAGE);
  uint64_t* full_bar = bars;        // [nstages <= 4]
  uint64_t* empty_bar = bars + 4;   // [4]
  uint64_t* tfull = bars + 8;       // [2]
  uint64_t* tempty = bars + 10;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.z * p.kb_per;
  const int kb1 = min(kb0 + p.kb_per, p.kb_total);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nstages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // two BN-column accumulators, double-buffered against the promotion
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * BN;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: per k-block the three planes of the A tile and of the B tile
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t bar = smem_u32(&full_bar[stage]);
        mbar_expect_tx(bar, STAGE);
        unsigned char* st = smem + (size_t)stage * STAGE;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          tma_load_2d(smem_u32(st + pl * GT_A_SUB), &tmapA, kb * GT_BK, (int)(pl * p.m_pad + m0), bar);
          tma_load_2d(smem_u32(st + 3 * GT_A_SUB + pl * B_SUB), &tmapB, kb * GT_BK, (int)(pl * p.n_pad + n0), bar);
        }
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer: one k-block = one fresh accumulator
      constexpr uint32_t idesc = umma_idesc_bf16(GT_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, acc = it & 1;
        mbar_wait(smem_u32(&tempty[acc]), ((uint32_t)(it >> 1) & 1u) ^ 1u);  // the promotion drained this accumulator's previous block
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(smem + (size_t)stage * STAGE);
        const uint32_t b_base = a_base + 3 * GT_A_SUB;
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        // smallest terms first: (A2,B2) (A1,B3) (A3,B1) (A1,B2) (A2,B1) (A1,B1)
        const int ia[6] = {1, 0, 2, 0, 1, 0};
        const int ib[6] = {1, 2, 0, 1, 0, 0};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
          for (int k = 0; k < GT_BK / 16; ++k) {
            tcgen05_mma_f16(d_tmem, umma_desc_sw128(a_base + ia[c] * GT_A_SUB + k * 32), umma_desc_sw128(b_base + ib[c] * B_SUB + k * 32),
                            idesc, (uint32_t)(c != 0 || k != 0));
          }
        }
        tcgen05_commit(smem_u32(&empty_bar[stage]));
        tcgen05_commit(smem_u32(&tfull[acc]));
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---- promotion + epilogue: thread = output row (TMEM lane); fp32 register accumulators, IEEE round-to-nearest adds
    const int q = warp & 3;
    float accr[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) accr[j] = 0.f;
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, acc = it & 1;
      mbar_wait(smem_u32(&tfull[acc]), (uint32_t)(it >> 1) & 1u);
      __syncwarp();
      tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_base + (uint32_t)c0, v);
        tmem_ld_wait_for(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) accr[c0 + j] = __fadd_rn(accr[c0 + j], __uint_as_float(v[j]));
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&tempty[acc]));
    }
    const long long row = m0 + q * 32 + lane;
    if (row < p.M) {
      float* out = p.C + (p.splits > 1 ? (long long)blockIdx.z * p.M * p.N : 0) + row * (p.splits > 1 ? p.N : p.ldc);
#pragma unroll
      for (int j = 0; j < BN; ++j) {
        const long long col = n0 + j;
        if (col < p.N) out[col] = accr[j];
      }
    }
  }
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long M, long long N, float* __restrict__ C, long long ldc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = part[i];
  for (int sp = 1; sp < splits; ++sp) s += part[(long long)sp * M * N + i];  // fixed order: deterministic
  C[(i / N) * ldc + (i % N)] = s;
}

struct GemmTcPlan {
  long long m_pad, n_pad, k_pad;
  int bn, kb_total, splits, kb_per, nstages;
  size_t off_a, off_b, off_part, total, smem;
};

static GemmTcPlan gemm_tc_plan(long long m, long long n, long long k) {
  GemmTcPlan P{};
  P.bn = n <= 64 ? 64 : 128;
  P.m_pad = cdiv(m, GT_BM) * GT_BM;
  P.n_pad = cdiv(n, P.bn) * P.bn;
  P.k_pad = cdiv(k, GT_BK) * GT_BK;
  P.kb_total = (int)(P.k_pad / GT_BK);
  const long long tiles = (P.m_pad / GT_BM) * (P.n_pad / P.bn);
  // K splits: fill the 148 SMs when there are few output tiles (e.g. H^T dE: 3 tiles, K = n_users), at least 4 k-blocks each;
  // the partial tiles are added by splitk_reduce_kernel in split order (deterministic)
  long long splits = 1;
  if (tiles < kNumSMs) splits = std::min<long long>(cdiv(kNumSMs, tiles), std::max<long long>(1, P.kb_total / 4));
  P.kb_per = (int)cdiv(P.kb_total, splits);
  P.splits = (int)cdiv(P.kb_total, P.kb_per);
  const int stage = 3 * GT_A_SUB + 3 * P.bn * GT_BK * 2;
  P.nstages = std::min(4, (int)((227 * 1024 - 2048) / stage));
  P.smem = 1024 + (size_t)P.nstages * stage + 128;
  size_t o = 0;
  P.off_a = o; o += ((size_t)3 * P.m_pad * P.k_pad * 2 + 1023) / 1024 * 1024;
  P.off_b = o; o += ((size_t)3 * P.n_pad * P.k_pad * 2 + 1023) / 1024 * 1024;
  P.off_part = o; o += P.splits > 1 ? (size_t)P.splits * m * n * 4 : 0;
  P.total = o + 2048;
  return P;
}

}  // namespace tmf

using namespace tmf;

extern "C" size_t tmf_gemm_tc_ws_bytes(int64_t m, int64_t n, int64_t k) {
  if (m <= 0 || n <= 0 || k <= 0) return 2048;
  return gemm_tc_plan(m, n, k).total;
}

extern "C" int tmf_gemm_tc(int32_t ta, int32_t tb, int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* B,
                           int64_t ldb, float* C, int64_t ldc, void* ws, size_t ws_bytes, tmf_stream_t stream) {
  TMF_REQUIRE(m >= 0 && n >= 0 && k >= 0 && A && B && C, "tmf_gemm_tc: bad arguments");
  if (m == 0 || n == 0) return TMF_OK;
  cudaStream_t st = as_stream(stream);
  if (k == 0) {
    TMF_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)n * 4, (size_t)m, st));
    return TMF_OK;
  }
  const GemmTcPlan P = gemm_tc_plan(m, n, k);
  TMF_REQUIRE(ws_bytes >= P.total, "tmf_gemm_tc: workspace too small");
  TMF_REQUIRE(3 * P.m_pad < (1ll << 31) && 3 * P.n_pad < (1ll << 31) && P.k_pad < (1ll << 31), "tmf_gemm_tc: sizes must fit int32 tile coordinates");
  unsigned char* w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* Ap = reinterpret_cast<__nv_bfloat16*>(w + P.off_a);
  __nv_bfloat16* Bp = reinterpret_cast<__nv_bfloat16*>(w + P.off_b);
  float* part = reinterpret_cast<float*>(w + P.off_part);
  // op(A) is [m, k]: stored [m, k] (ta = 0: K contiguous) or [k, m] (ta = 1: transpose on the way in); op(B)^T is [n, k]:
  // stored [n, k] (tb = 1) or [k, n] (tb = 0: transpose)
  {
    const long long tiles = (P.m_pad / 32) * (P.k_pad / 32);
    split3_pack_kernel<<<(unsigned)cdiv(tiles, 8), 256, 0, st>>>(A, ta ? k : m, ta ? m : k, lda, ta ? 1 : 0, Ap, P.m_pad, P.k_pad);
  }
  {
    const long long tiles = (P.n_pad / 32) * (P.k_pad / 32);
    split3_pack_kernel<<<(unsigned)cdiv(tiles, 8), 256, 0, st>>>(B, tb ? n : k, tb ? k : n, ldb, tb ? 0 : 1, Bp, P.n_pad, P.k_pad);
  }
  TMF_LAUNCH_CHECK();
  CUtensorMap tmA, tmB;
  int rc = make_tmap_k64(&tmA, Ap, 3 * P.m_pad, P.k_pad, GT_BM);
  if (rc) return rc;
  rc = make_tmap_k64(&tmB, Bp, 3 * P.n_pad, P.k_pad, P.bn);
  if (rc) return rc;
  GemmTcParams p{};
  p.M = m; p.N = n; p.m_pad = P.m_pad; p.n_pad = P.n_pad; p.kb_total = P.kb_total; p.splits = P.splits; p.kb_per = P.kb_per;
  p.nstages = P.nstages; p.C = P.splits > 1 ? part : C; p.ldc = ldc;
  dim3 grid((unsigned)(P.n_pad / P.bn), (unsigned)(P.m_pad / GT_BM), (unsigned)P.splits);
  TMF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "tmf_gemm_tc: too many row tiles / splits for one launch");
  if (P.bn == 64) {
    TMF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    gemm_tc_kernel<64><<<grid, GT_THREADS, P.smem, st>>>(tmA, tmB, p);
  } else {
    TMF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
    gemm_tc_kernel<128><<<grid, GT_THREADS, P.smem, st>>>(tmA, tmB, p);
  }
  TMF_LAUNCH_CHECK();
  if (P.splits > 1) {
    splitk_reduce_kernel<<<(unsigned)cdiv(m * n, 256), 256, 0, st>>>(part, P.splits, m, n, C, ldc);
    TMF_LAUNCH_CHECK();
  }
  return TMF_OK;
}
