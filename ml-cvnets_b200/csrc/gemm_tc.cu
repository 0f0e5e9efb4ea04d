// tcgen05 / TMEM / TMA pointwise-conv GEMM (sm_100a): every load mode, STORE / residual / SiLU-backward / GroupNorm-backward epilogues.
//
//   C[M,N] = epi( A[M,K] * W[N,K]^T + bias )          same contract as the mma.sync kernel in gemm.cu
//
// Warp-specialised, one persistent CTA per SM:
//   warp 0    TMA producer : ring of A stages [128 pixels x 32 k] (cp.async.bulk.tensor.2d, 64-byte swizzle) across ALL tiles
//   warp 1    MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16; the accumulator lives in TMEM,
//                            double buffered (2 x 128 columns), completion is signalled with tcgen05.commit -> mbarrier
//   warps 2-17 epilogue    : tcgen05.ld (32 lanes x 32 columns), bias / residual / activation-backward / BatchNorm statistics,
//                            bf16 staging tile in smem, 16-byte row-contiguous stores
//   warps 18-21 transform  : (layers with a prologue) apply the producer's BN(+SiLU) / GroupNorm / BN-backward to the landed A
//                            stage in place, fence.proxy.async, then hand the stage to the MMA warp through a second mbarrier
// The product is computed TRANSPOSED, D[channel, pixel] = W[channel, :] . A[pixel, :], i.e. the weight panel is the UMMA
// "A" operand (M = 128 output channels = TMEM lanes) and the activation tile the "B" operand (N = 128 pixels = TMEM columns).
// Each epilogue thread then owns ONE output channel: bias is a scalar, the per-channel BatchNorm sums are thread-local (no
// shuffles, no smem atomics) and are flushed with one fp64 atomic per channel per CTA.  The epilogue of tile j overlaps the
// MMAs of tile j+1 and the TMA loads of tiles j+2...
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int TC_BM = 128;      // pixels per tile  (UMMA N)
constexpr int TC_BN = 128;      // channels per tile (UMMA M)
constexpr int TC_BK = 32;       // k per stage (64-byte rows)
constexpr int TC_STAGE = TC_BM * TC_BK * 2;   // 8 KB
constexpr int TC_WBLK = TC_BN * TC_BK * 2;    // 8 KB per k-block of the weight panel
constexpr int TC_LDO = TC_BN + 8;             // bf16 staging row stride (elements)
constexpr int TC_EPI_WARPS = 16;            // four warps per TMEM lane quadrant, each draining a quarter (32) of the pixel columns: the
                                            // epilogue, not HBM, bounded round 1's kernel (2 warps / scheduler, ~3000 cycles per tile)
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;
constexpr int TC_XF_THREADS_MAX = 256;      // transform warps (only launched for layers with a prologue).  Each warp is a latency-bound
                                            // chain (LDS -> convert -> FMA -> MUFU -> pack -> STS): with 4 warps the MMA issuer spent most of its
                                            // time waiting for transformed stages (ncu source view, round 2), so the prologue layers run 8
                                            // (same-box A/B: AFF_SILU 100.9 -> 85.9 us, GN 48.3 -> 43.4 us); the BNB prologue (two operand
                                            // tiles per stage, SiLU-backward epilogue) was 4 % faster with 4 and keeps them
template <int AMODE>
struct XfCfg {
  static constexpr int THREADS = (AMODE == CVB_A_RAW) ? 0 : (AMODE == CVB_A_BNB ? 128 : TC_XF_THREADS_MAX);
  static constexpr int IT = THREADS ? (TC_BM * 4) / THREADS : 1;  // 16-byte chunks of a [128 x 32] bf16 stage per transform thread
};
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_TMEM_COLS = 256;             // 2 accumulators x 128 fp32 columns

enum { TEPI_STORE = 0, TEPI_STORE_R = 1, TEPI_SILU_BWD = 2, TEPI_GN_BWD = 3 };

__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  // K-major operand, 64-byte swizzle: rows of 64 B, 8-row groups 512 B apart (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (16 B; unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(512 >> 4) << 32;           // stride byte offset = 512 B bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version 1 (Blackwell)
  d |= (uint64_t)4 << 61;                    // layout type: SWIZZLE_64B
  return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128 (>>3), M = 128 (>>4)
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 3) << 17) | ((uint32_t)(TC_BN >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }

// WRES: the weight panel [128 ch, K] stays resident in smem (loaded once); otherwise (large K) its k-blocks stream through the
// ring next to the activation k-blocks (they are L2 hits: every CTA of an N tile reads the same panel).
template <int AMODE, int EPI, bool WRES>
__global__ void __launch_bounds__(TC_THREADS + XfCfg<AMODE>::THREADS, 1)
    pw_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW,
                      const cvb_gemm_args p, int NST) {
  constexpr bool XF = (AMODE != CVB_A_RAW);     // has transform warps
  constexpr bool TWO_A = (AMODE == CVB_A_BNB);  // BN-backward prologue streams two tensors
  constexpr bool HAS_P = (AMODE == CVB_A_AFF || AMODE == CVB_A_AFF_SILU || AMODE == CVB_A_GN || AMODE == CVB_A_BNB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * TC_BN;
  const int KT = (p.K + TC_BK - 1) / TC_BK;
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int my_tiles = (m_tiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
  const int total = my_tiles * KT;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = (TWO_A ? 2 : 1) * TC_STAGE;
  constexpr int RING_STAGE = A_BYTES + (WRES ? 0 : TC_WBLK);  // [A | A2 (BNB) | W block (streaming)]
  uint8_t* sW = smem;                               // resident weight panel: KT blocks [128 ch][32 k] (WRES only)
  uint8_t* sA = sW + (WRES ? KT * TC_WBLK : 0);     // ring
  uint8_t* sO = sA + NST * RING_STAGE;              // bf16 [128 pix][TC_LDO] staging (aux in / result out)
  float* sP = reinterpret_cast<float*>(sO + TC_BM * TC_LDO * 2);  // prologue parameters [3][Kpad]
  __shared__ __align__(8) uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], ready[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t wbar, tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ double s_samp[2][128];

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&ready[i], XfCfg<AMODE>::THREADS ? XfCfg<AMODE>::THREADS / 32 : 1); }
    mbar_init(&wbar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TC_EPI_WARPS); }
    fence_mbar_init();
  }
  if (tid < 128) { s_samp[0][tid] = 0.0; s_samp[1][tid] = 0.0; }
  if (warp == 1) {  // TMEM allocation (whole warp), address published through smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // barriers, TMEM allocation and CTA scheduling overlapped the previous kernel's tail; data accesses start here
  pdl_trigger();

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      if (WRES) {
        mbar_expect_tx(&wbar, (uint32_t)KT * TC_WBLK);
        for (int kt = 0; kt < KT; ++kt) tma_load_2d(sW + kt * TC_WBLK, &tmW, &wbar, kt * TC_BK, n0);
      }
      for (int it = 0; it < total; ++it) {
        const int stage = it % NST;
        if (it >= NST) mbar_wait(&empty[stage], ((it / NST) - 1) & 1);  // MMAs that read this slot have completed
        const int j = it / KT, kt = it - j * KT;
        const int m0 = ((int)blockIdx.y + j * (int)gridDim.y) * TC_BM;
        mbar_expect_tx(&full[stage], RING_STAGE);
        tma_load_2d(sA + stage * RING_STAGE, &tmA, &full[stage], kt * TC_BK, m0);
        if (TWO_A) tma_load_2d(sA + stage * RING_STAGE + TC_STAGE, &tmA2, &full[stage], kt * TC_BK, m0);
        if (!WRES) tma_load_2d(sA + stage * RING_STAGE + A_BYTES, &tmW, &full[stage], kt * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      if (WRES) mbar_wait(&wbar, 0);
      int it = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const int buf = j & 1;
        if (j >= 2) mbar_wait(&tempty[buf], ((j >> 1) - 1) & 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TC_BM);
        for (int kt = 0; kt < KT; ++kt, ++it) {
          const int stage = it % NST;
          mbar_wait(XF ? &ready[stage] : &full[stage], (it / NST) & 1);  // landed (and transformed in place)
          tc_fence_after();
          const uint32_t aa = smem_u32(sA + stage * RING_STAGE);
          const uint32_t wa = WRES ? smem_u32(sW + kt * TC_WBLK) : aa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_f16(tmem_d, umma_desc_sw64(wa + k * 32), umma_desc_sw64(aa + k * 32), (kt | k) ? 1u : 0u);
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
        }
        umma_commit(&tfull[buf]);      // accumulator complete
      }
    }
  } else if (warp >= 2 + TC_EPI_WARPS) {
    // ===================================================== transform warps: producer's normalisation / activation, in place
    if (XF) {
      constexpr int TC_XF_THREADS = XfCfg<AMODE>::THREADS > 0 ? XfCfg<AMODE>::THREADS : 128;
      constexpr int TC_XF_IT = XfCfg<AMODE>::IT;
      const int tt = tid - TC_THREADS;  // 0..TC_XF_THREADS-1
      const int Kpad = KT * TC_BK;
      if (HAS_P) {
        for (int k = tt; k < Kpad; k += TC_XF_THREADS) {
          const bool ok = k < p.K;
          sP[k] = ok ? p.a_p0[k] : 0.f;
          sP[Kpad + k] = ok ? p.a_p1[k] : 0.f;
          if (AMODE == CVB_A_BNB) sP[2 * Kpad + k] = ok ? p.a_p2[k] : 0.f;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(TC_XF_THREADS) : "memory");
      }
      float tmu[TC_XF_IT], trs[TC_XF_IT];
#pragma unroll
      for (int i = 0; i < TC_XF_IT; ++i) { tmu[i] = 0.f; trs[i] = 1.f; }
      for (int it = 0; it < total; ++it) {
        const int stage = it % NST;
        const int j = it / KT, kt = it - j * KT;
        const int m0 = ((int)blockIdx.y + j * (int)gridDim.y) * TC_BM;
        const int k0 = kt * TC_BK;
        if (AMODE == CVB_A_GN && kt == 0) {
#pragma unroll
          for (int i = 0; i < TC_XF_IT; ++i) {
            const int m = m0 + (tt >> 2) + i * (TC_XF_THREADS / 4);
            const int b = (m < p.M ? m : p.M - 1) / p.rows_per_sample;
            tmu[i] = __ldg(p.row_mean + b);
            trs[i] = __ldg(p.row_rstd + b);
          }
        }
        mbar_wait(&full[stage], (it / NST) & 1);
        uint8_t* st = sA + stage * RING_STAGE;
#pragma unroll
        for (int i = 0; i < TC_XF_IT; ++i) {
          const int c = tt + i * TC_XF_THREADS;
          const int row = c >> 2, ch = c & 3;
          const int k = k0 + ch * 8;
          const uint32_t off = (uint32_t)(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));  // 64-byte swizzle (TMA == UMMA layout)
          uint4* pa = reinterpret_cast<uint4*>(st + off);
          float f[8], q0[8], q1[8];
          unpack8(*pa, f);
          if (HAS_P) {
            *reinterpret_cast<float4*>(q0) = *reinterpret_cast<const float4*>(sP + k);
            *reinterpret_cast<float4*>(q0 + 4) = *reinterpret_cast<const float4*>(sP + k + 4);
            *reinterpret_cast<float4*>(q1) = *reinterpret_cast<const float4*>(sP + Kpad + k);
            *reinterpret_cast<float4*>(q1 + 4) = *reinterpret_cast<const float4*>(sP + Kpad + k + 4);
          }
          if (AMODE == CVB_A_AFF) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(q0[e], f[e], q1[e]);
          } else if (AMODE == CVB_A_AFF_SILU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = silu_f(fmaf(q0[e], f[e], q1[e]));
          } else if (AMODE == CVB_A_SILU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = silu_f(f[e]);
          } else if (AMODE == CVB_A_GN) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf((f[e] - tmu[i]) * trs[i], q0[e], q1[e]);
          } else if (AMODE == CVB_A_BNB) {
            float y[8], q2[8];
            unpack8(*reinterpret_cast<const uint4*>(st + TC_STAGE + off), y);
            *reinterpret_cast<float4*>(q2) = *reinterpret_cast<const float4*>(sP + 2 * Kpad + k);
            *reinterpret_cast<float4*>(q2 + 4) = *reinterpret_cast<const float4*>(sP + 2 * Kpad + k + 4);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(q0[e], f[e], fmaf(q1[e], y[e], q2[e]));
          }
          // rows beyond M must stay exactly zero (their accumulators would otherwise pollute the statistics)
          *pa = (m0 + row < p.M) ? pack8(f) : make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[stage]);
      }
    }
  } else {
    // ===================================================== epilogue warps (threads 64..319): one output channel per thread
    const int et = tid - 64;                 // 0..TC_EPI_THREADS-1
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int cquart = (warp - 2) >> 2;      // which quarter (32) of the pixel columns this warp drains
    const int ch_local = quad * 32 + lane;   // output channel within the tile == TMEM lane
    const int ch = n0 + ch_local;
    const bool ch_ok = ch < p.N;
    const float bias = (ch_ok && p.bias) ? __ldg(p.bias + ch) : 0.f;
    const float ep0 = ((EPI == TEPI_SILU_BWD || EPI == TEPI_GN_BWD) && ch_ok && p.e_p0) ? __ldg(p.e_p0 + ch) : 1.f;
    const float ep1 = (EPI == TEPI_SILU_BWD && ch_ok && p.e_p1) ? __ldg(p.e_p1 + ch) : 0.f;
    constexpr bool has_aux = (EPI != TEPI_STORE);
    const bf16* __restrict__ AUX = static_cast<const bf16*>(EPI == TEPI_STORE_R ? p.R : p.Y);
    const int ldaux = EPI == TEPI_STORE_R ? p.ldr : p.ldy;
    const bool want_samp = (p.samp_sum != nullptr) && (EPI != TEPI_GN_BWD);  // GN_BWD: the sample sums come from the workspace finalize
    const bool lin_bwd = (p.e_mode == CVB_E_LIN_BWD);  // SiLU-backward epilogue without the activation factor (BatchNorm with no act)
    const int rps = p.rows_per_sample > 0 ? p.rows_per_sample : 1;
    float cs = 0.f, cq = 0.f;  // this channel's statistics over all tiles of the CTA
    float2 cs2 = make_float2(0.f, 0.f), cq2 = make_float2(0.f, 0.f);  // hot-path partials (even / odd pixel columns), folded in at the end
    bf16* __restrict__ Cg = static_cast<bf16*>(p.C);
    constexpr int CGS = TC_BN / 8;

    auto issue_aux = [&](int j) {
      const int m0 = ((int)blockIdx.y + j * (int)gridDim.y) * TC_BM;
      for (int c = et; c < TC_BM * CGS; c += TC_EPI_THREADS) {
        const int row = c / CGS, cgc = c % CGS;
        const int m = m0 + row, n = n0 + cgc * 8;
        const bool ok = (m < p.M) && (n < p.N);
        cp_async16(smem_u32(sO + row * (TC_LDO * 2) + cgc * 16), AUX + (ok ? (size_t)m * ldaux + n : 0), ok);
      }
      cp_async_commit();
    };
    if (has_aux && my_tiles > 0) issue_aux(0);

    for (int j = 0; j < my_tiles; ++j) {
      const int buf = j & 1;
      const int m0 = ((int)blockIdx.y + j * (int)gridDim.y) * TC_BM;
      if (has_aux) {
        cp_async_wait<0>();
        epi_bar_sync();  // aux tile visible to all epilogue threads
      }
      mbar_wait(&tfull[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * TC_BM);
      const bool full_tile = (m0 + TC_BM <= p.M);  // rows >= M have zero A rows; only their bias must be masked (last tile)
      {
        const int cc = cquart;
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        uint16_t* so = reinterpret_cast<uint16_t*>(sO) + (cc * 32) * TC_LDO + ch_local;
        if (full_tile && !has_aux) {
          // hot path (conv -> BN statistics): ~5 instructions per value.  The BatchNorm sums are taken from the fp32 values
          // (before bf16 rounding): the rounding error averages out over the >= 128 pixels of the tile.
          // packed fp32 pairs (adjacent pixel columns are adjacent registers of the tcgen05.ld result): 7 instructions per 2 values
          const float2 b2 = make_float2(bias, bias);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 v = fadd2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), b2);
            const uint32_t pk = pack_bf162(v.x, v.y);
            so[i * TC_LDO] = (uint16_t)pk;
            so[(i + 1) * TC_LDO] = (uint16_t)(pk >> 16);
            cs2 = fadd2(cs2, v);
            cq2 = ffma2(v, v, cq2);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(r[i]) + ((full_tile || m0 + cc * 32 + i < p.M) ? bias : 0.f);
            float y = 0.f;
            if (has_aux) y = __bfloat162float(reinterpret_cast<const bf16*>(so)[i * TC_LDO]);
            if (EPI == TEPI_STORE_R) v += y;
            if (EPI == TEPI_SILU_BWD && !lin_bwd) v *= silu_grad_f(fmaf(ep0, y, ep1));
            if (EPI == TEPI_GN_BWD) {
              // GroupNorm backward, phase 1 in sum form: per (sample, channel) A = sum v, Bx = sum v*x (raw x); everything else
              // (dgamma, dbeta, per-sample sums of g and g*xhat) is linear in A and Bx and is derived by the finalize kernel
              cs += v;
              cq = fmaf(v, y, cq);
              reinterpret_cast<bf16*>(so)[i * TC_LDO] = __float2bfloat16_rn(v * ep0);
            } else {
              const bf16 vb = __float2bfloat16_rn(v);
              reinterpret_cast<bf16*>(so)[i * TC_LDO] = vb;
              const float vr = __bfloat162float(vb);  // statistics of the STORED values
              cs += vr;
              cq = fmaf(vr, EPI == TEPI_SILU_BWD ? y : vr, cq);
            }
          }
        }
      }
      if (EPI == TEPI_GN_BWD) {  // this thread's 32 pixel columns never straddle samples (rows_per_sample % 64 == 0, checked on the host)
        const int mh = m0 + cquart * 32;
        if (ch_ok && mh < p.M) {
          const int nsamples = (p.M + rps - 1) / rps;
          double* wsA = p.gn_ws + (size_t)(mh / rps) * p.N + ch;
          atomicAdd(wsA, (double)cs);
          atomicAdd(wsA + (size_t)nsamples * p.N, (double)cq);
        }
        cs = 0.f;
        cq = 0.f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);  // one arrival per epilogue warp releases the accumulator
      epi_bar_sync();                            // staged tile complete
      const int first_sample = m0 / rps;
      for (int c = et; c < TC_BM * CGS; c += TC_EPI_THREADS) {
        const int row = c / CGS, cgc = c % CGS;
        const int m = m0 + row, n = n0 + cgc * 8;
        const uint4 u = *reinterpret_cast<const uint4*>(sO + row * (TC_LDO * 2) + cgc * 16);
        if (m < p.M && n < p.N) stg16(Cg + (size_t)m * p.ldc + n, u);
        if (want_samp) {
          float f[8];
          unpack8(u, f);
          float sv = 0.f, sq = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) { sv += f[e]; sq = fmaf(f[e], f[e], sq); }
#pragma unroll
          for (int o = CGS / 2; o > 0; o >>= 1) {
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          if (cgc == 0 && m < p.M) {
            atomicAdd(&s_samp[0][m / rps - first_sample], (double)sv);
            atomicAdd(&s_samp[1][m / rps - first_sample], (double)sq);
          }
        }
      }
      epi_bar_sync();  // staging tile free again (and s_samp complete)
      if (want_samp) {
        const int mlast = min(m0 + TC_BM, p.M) - 1;
        const int nsamp = mlast / rps - first_sample + 1;
        if (et < nsamp) {
          atomicAdd(p.samp_sum + first_sample + et, s_samp[0][et]);
          atomicAdd(p.samp_sq + first_sample + et, s_samp[1][et]);
          s_samp[0][et] = 0.0;
          s_samp[1][et] = 0.0;
        }
      }
      if (has_aux && j + 1 < my_tiles) issue_aux(j + 1);
    }
    cs += cs2.x + cs2.y;
    cq += cq2.x + cq2.y;
    if (EPI != TEPI_GN_BWD && p.col_sum && ch_ok) {
      atomicAdd(p.col_sum + ch, (double)cs);
      atomicAdd(p.col_sq + ch, (double)cq);
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

template <int AMODE, int EPI, bool WRES>
int launch_tc_impl(const cvb_gemm_args& a, cudaStream_t st, size_t fixed, int stage_bytes) {
  const size_t budget = (size_t)216 * 1024;
  int nst = (int)((budget - fixed) / stage_bytes);
  if (nst > TC_MAX_STAGES) nst = TC_MAX_STAGES;
  const size_t smem = fixed + (size_t)nst * stage_bytes;
  static bool attr = false;
  if (!attr) {
    CVB_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<AMODE, EPI, WRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    attr = true;
  }
  const int n_tiles = (a.N + TC_BN - 1) / TC_BN, m_tiles = (a.M + TC_BM - 1) / TC_BM;
  // one persistent CTA per SM, never more CTAs than SMs: rounding UP (3 x 50 = 150 CTAs on 148 SMs) put two CTAs into a second wave and doubled
  // the time of every N = 264 / 392 / 520 / 768-wide late-stage layer (profiles/r2_step_launches_v2_lazybn_dram.csv)
  int gy = cvb_num_sms() / n_tiles;
  if (gy > m_tiles) gy = m_tiles;
  if (gy < 1) gy = 1;
  CUtensorMap tmA, tmA2, tmW;
  if (cvb_make_tmap_2d_k32(&tmA, a.A, a.M, a.K, a.lda, TC_BM)) return 1;
  if (cvb_make_tmap_2d_k32(&tmA2, AMODE == CVB_A_BNB ? a.A2 : a.A, a.M, a.K, AMODE == CVB_A_BNB ? a.lda2 : a.lda, TC_BM)) return 1;
  if (cvb_make_tmap_2d_k32(&tmW, a.W, a.N, a.K, a.ldw, TC_BN)) return 1;
  dim3 grid(n_tiles, gy);
  const int threads = TC_THREADS + XfCfg<AMODE>::THREADS;
  CVB_CUDA(cvb_launch(pw_gemm_tc_kernel<AMODE, EPI, WRES>, grid, threads, smem, st, tmA, tmA2, tmW, a, nst));
  CVB_LAUNCH_CHECK();
  return 0;
}

template <int AMODE, int EPI>
int launch_tc(const cvb_gemm_args& a, cudaStream_t st) {
  const int KT = (a.K + TC_BK - 1) / TC_BK;
  const int nvec = (AMODE == CVB_A_AFF || AMODE == CVB_A_AFF_SILU || AMODE == CVB_A_GN) ? 2 : (AMODE == CVB_A_BNB ? 3 : 0);
  const int a_bytes = (AMODE == CVB_A_BNB ? 2 : 1) * TC_STAGE;
  const size_t stagebuf = (size_t)TC_BM * TC_LDO * 2 + (size_t)nvec * KT * TC_BK * 4 + 1024;
  const size_t panel = (size_t)KT * TC_WBLK;
  if (panel + stagebuf + 6 * (size_t)a_bytes <= (size_t)216 * 1024) return launch_tc_impl<AMODE, EPI, true>(a, st, panel + stagebuf, a_bytes);
  return launch_tc_impl<AMODE, EPI, false>(a, st, stagebuf, a_bytes + TC_WBLK);  // large K: weight k-blocks ride the ring
}

// GroupNorm backward, phase 1 finalize: from A[b,c] = sum_m v, Bx[b,c] = sum_m v*x over the pixels of sample b
//   dbeta[c] += A;  dgamma[c] += t,  t = rstd_b (Bx - mean_b A) = sum_m v*xhat;   sum g = sum_c gamma_c A;   sum g*xhat = sum_c gamma_c t
__global__ void __launch_bounds__(128) gn_bwd_ws_finalize_kernel(const double* __restrict__ ws, const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, const float* __restrict__ gamma, int B, int N,
                                                                  double* col_sum, double* col_sq, double* samp_sum, double* samp_sq) {
  pdl_wait();
  pdl_trigger();
  __shared__ double s_red[2][4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double mu = (double)mean[b], rs = (double)rstd[b];
  double sg = 0.0, sgx = 0.0;
  for (int c = tid; c < N; c += 128) {
    const double A = ws[(size_t)b * N + c], Bx = ws[((size_t)B + b) * N + c];
    const double t = rs * (Bx - mu * A);
    const double gm = gamma ? (double)gamma[c] : 1.0;
    if (col_sum) { atomicAdd(col_sum + c, A); atomicAdd(col_sq + c, t); }
    sg += gm * A;
    sgx += gm * t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
  }
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = sg; s_red[1][tid >> 5] = sgx; }
  __syncthreads();
  if (tid == 0 && samp_sum) {
    atomicAdd(samp_sum + b, s_red[0][0] + s_red[0][1] + s_red[0][2] + s_red[0][3]);
    atomicAdd(samp_sq + b, s_red[1][0] + s_red[1][1] + s_red[1][2] + s_red[1][3]);
  }
}

template <int AMODE>
int launch_tc_gn_bwd(const cvb_gemm_args& a, cudaStream_t st) {
  int rc = launch_tc<AMODE, TEPI_GN_BWD>(a, st);
  if (rc != 0) return rc;
  const int B = (a.M + a.rows_per_sample - 1) / a.rows_per_sample;
  CVB_CUDA(cvb_launch(gn_bwd_ws_finalize_kernel, B, 128, 0, st, static_cast<const double*>(a.gn_ws), a.row_mean, a.row_rstd, a.e_p0, B, a.N, a.col_sum,
                      a.col_sq, a.samp_sum, a.samp_sq));
  CVB_LAUNCH_CHECK();
  return 0;
}

template <int AMODE>
int dispatch_tc_epi(const cvb_gemm_args& a, cudaStream_t st) {
  if (a.e_mode == CVB_E_GN_BWD && a.gn_ws && a.rows_per_sample % (TC_BM / 2) == 0 && !a.bias && (AMODE == CVB_A_RAW || AMODE == CVB_A_BNB))
    return launch_tc_gn_bwd<AMODE == CVB_A_BNB ? CVB_A_BNB : CVB_A_RAW>(a, st);
  if (a.e_mode == CVB_E_STORE) return a.R ? launch_tc<AMODE, TEPI_STORE_R>(a, st) : launch_tc<AMODE, TEPI_STORE>(a, st);
  if ((a.e_mode == CVB_E_SILU_BWD || a.e_mode == CVB_E_LIN_BWD) && (AMODE == CVB_A_RAW || AMODE == CVB_A_BNB)) return launch_tc<AMODE == CVB_A_BNB ? CVB_A_BNB : CVB_A_RAW, TEPI_SILU_BWD>(a, st);
  return -1;
}

}  // namespace

// Returns -1 when the shape / mode is not handled by the tcgen05 kernel (caller uses the mma.sync kernel), 0 on success, > 0 on error.
int cvb_pw_gemm_tc(const cvb_gemm_args& a, cudaStream_t st) {
  // every epilogue thread owns one of 128 output channels: narrow layers would idle most of them -> mma.sync kernel
  static const int min_n = [] { const char* e = getenv("CVB_TC_MIN_N"); return e ? atoi(e) : 96; }();  // diagnostics: route narrower layers here
  if (a.N < min_n || (a.N >= 96 && a.N % 128 != 0 && a.N % 128 < 64 && a.N < 256)) return -1;  // N = 64 on this kernel measured slower than mma.sync (round 1)
  switch (a.a_mode) {
    case CVB_A_RAW: return dispatch_tc_epi<CVB_A_RAW>(a, st);
    case CVB_A_AFF: return dispatch_tc_epi<CVB_A_AFF>(a, st);
    case CVB_A_AFF_SILU: return dispatch_tc_epi<CVB_A_AFF_SILU>(a, st);
    case CVB_A_SILU: return dispatch_tc_epi<CVB_A_SILU>(a, st);
    case CVB_A_GN: return dispatch_tc_epi<CVB_A_GN>(a, st);
    case CVB_A_BNB: return dispatch_tc_epi<CVB_A_BNB>(a, st);
    default: return -1;
  }
}
