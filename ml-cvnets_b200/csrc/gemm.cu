// Pointwise-conv / linear GEMMs for the MobileViTv2 hot path (sm_100a).
//
//   cvb_pw_gemm : C[M,N] = epi( load(A)[M,K] * W[N,K]^T + bias )      forward and input-gradient GEMMs
//   cvb_pw_wgrad: dW[N,K] += sum_m load(G)[m,n] * load(A)[m,k]        weight-gradient GEMM (reduction over pixels)
//
// Every layer here is HBM-bound (K,N <= 768: <= 170 FLOP/B, B200 ridge ~250 FLOP/B), so the design goal is
// "read each activation once, write each activation once": the producer's BatchNorm/SiLU/GroupNorm is applied to the
// A fragments in registers between ldmatrix and mma (no extra pass, no extra smem traffic), and bias / activation /
// residual / BN-statistics / GN-statistics / activation-backward are applied in the epilogue on a smem-staged tile so
// that all global traffic is 16-byte, row-contiguous.  Tensor-core path: mma.sync.m16n8k16 bf16 -> fp32 (register
// prologue is what makes the fusion free); operands arrive through a 4-stage cp.async ring with XOR-swizzled smem.
#include "common.cuh"

namespace {

constexpr int BM = 128;       // CTA tile rows (pixels)
constexpr int BK = 32;        // K step
constexpr int NTHREADS = 256;

__device__ __forceinline__ uint32_t swz64(int row, int ch) {  // 64-byte rows, 4 x 16B chunks
  return static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
}

template <int AMODE>
struct GemmCfg {
  // the BN-backward prologue streams two A tensors: one stage less keeps two CTAs per SM
  static constexpr int kStages = (AMODE == CVB_A_BNB) ? 3 : 4;
};

template <int WM, int AMODE>
__global__ void __launch_bounds__(NTHREADS, 2) pw_gemm_kernel(const cvb_gemm_args p) {
  constexpr int WARPS_M = BM / WM;
  constexpr int WARPS_N = 8 / WARPS_M;
  constexpr int BN = WARPS_N * 32;
  constexpr int MI = WM / 16;
  constexpr bool TWO_A = (AMODE == CVB_A_BNB);
  constexpr int NST = GemmCfg<AMODE>::kStages;
  constexpr int A_STAGE = BM * BK * 2;
  constexpr int B_STAGE = BN * BK * 2;
  constexpr int LDC_S = BN + 4;      // fp32 staging row stride
  constexpr int HALF = BM / 2;       // the epilogue stages the tile in two 64-row halves
  constexpr bool HAS_P = (AMODE == CVB_A_AFF || AMODE == CVB_A_AFF_SILU || AMODE == CVB_A_GN || AMODE == CVB_A_BNB);

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sA2 = smem + NST * A_STAGE;
  uint8_t* sB = smem + (TWO_A ? 2 : 1) * NST * A_STAGE;
  constexpr int PIPE_BYTES = (TWO_A ? 2 : 1) * NST * A_STAGE + NST * B_STAGE;
  float* sC = reinterpret_cast<float*>(smem + PIPE_BYTES);                      // NOT overlaid: loads of the next tiles stay in flight
  float* sP = reinterpret_cast<float*>(smem + PIPE_BYTES + HALF * LDC_S * 4);   // prologue parameters
  __shared__ float s_col[2][128];
  __shared__ double s_samp[2][128];  // fp64: cross-thread order must not change the GroupNorm statistics

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / WARPS_N) * WM;
  const int wn0 = (warp % WARPS_N) * 32;
  const int n0 = blockIdx.x * BN;
  const int KT = (p.K + BK - 1) / BK;
  const int Kpad = KT * BK;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int my_tiles = (m_tiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
  const int total = my_tiles * KT;  // flattened (tile, k-tile) iterations of this CTA

  if (tid < 128) { s_col[0][tid] = 0.f; s_col[1][tid] = 0.f; s_samp[0][tid] = 0.0; s_samp[1][tid] = 0.0; }
  if (HAS_P) {  // per-K prologue parameters -> smem (zero padded so that the K tail transforms to finite values)
    for (int k = tid; k < Kpad; k += NTHREADS) {
      bool ok = k < p.K;
      sP[k] = ok ? p.a_p0[k] : 0.f;
      sP[Kpad + k] = ok ? p.a_p1[k] : 0.f;
      if (AMODE == CVB_A_BNB) sP[2 * Kpad + k] = ok ? p.a_p2[k] : 0.f;
    }
  }

  const bf16* __restrict__ A = static_cast<const bf16*>(p.A);
  const bf16* __restrict__ A2 = static_cast<const bf16*>(p.A2);
  const bf16* __restrict__ Wg = static_cast<const bf16*>(p.W);

  // epilogue thread mapping (fixed per thread across tiles): 8 consecutive columns of one row
  constexpr int CGS = BN / 8;
  constexpr int ROWS_PER_PASS = NTHREADS / CGS;
  const int cg = tid % CGS, r0 = tid / CGS;
  const int nc = n0 + cg * 8;
  const bool col_ok = nc < p.N;
  const int emode = p.e_mode;
  const bool want_col = p.col_sum != nullptr;
  const bool want_samp = p.samp_sum != nullptr;
  const int rps = p.rows_per_sample > 0 ? p.rows_per_sample : 1;
  const bf16* __restrict__ Yg = static_cast<const bf16*>(p.Y);
  const bf16* __restrict__ Rg = static_cast<const bf16*>(p.R);
  float cs[8], cq[8];  // per-column statistics, accumulated over all tiles of this CTA, flushed once
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }

  // one ring for the whole CTA lifetime: iteration `it` = (tile it / KT, k-tile it % KT)
  auto issue = [&](int it) {
    const int stage = it % NST;
    const int j = it / KT, kt = it - j * KT;
    const int m0i = ((int)blockIdx.y + j * (int)gridDim.y) * BM;
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int c = tid + i * NTHREADS;
      int row = c >> 2, ch = c & 3;
      int m = m0i + row, k = k0 + ch * 8;
      bool ok = (m < p.M) && (k < p.K);
      cp_async16(smem_u32(sA + stage * A_STAGE) + swz64(row, ch), A + (ok ? (size_t)m * p.lda + k : 0), ok);
      if (TWO_A) cp_async16(smem_u32(sA2 + stage * A_STAGE) + swz64(row, ch), A2 + (ok ? (size_t)m * p.lda2 + k : 0), ok);
    }
    for (int c = tid; c < BN * 4; c += NTHREADS) {
      int row = c >> 2, ch = c & 3;
      int n = n0 + row, k = k0 + ch * 8;
      bool ok = (n < p.N) && (k < p.K);
      cp_async16(smem_u32(sB + stage * B_STAGE) + swz64(row, ch), Wg + (ok ? (size_t)n * p.ldw + k : 0), ok);
    }
  };

#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < total) issue(s);
    cp_async_commit();
  }

  int it = 0;
  for (int jt = 0; jt < my_tiles; ++jt) {
    const int m0 = ((int)blockIdx.y + jt * (int)gridDim.y) * BM;
    // GroupNorm prologue: per-row statistics of the rows this thread's fragments touch
    float rmean[MI][2], rrstd[MI][2];
    if (AMODE == CVB_A_GN) {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int m = m0 + wm0 + mi * 16 + (lane >> 2) + h * 8;
          int b = (m < p.M ? m : p.M - 1) / p.rows_per_sample;
          rmean[mi][h] = p.row_mean[b];
          rrstd[mi][h] = p.row_rstd[b];
        }
    }
    float acc[MI][4][4];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;

    for (int kt = 0; kt < KT; ++kt, ++it) {
      cp_async_wait<NST - 2>();
      __syncthreads();
      {
        int nxt = it + NST - 1;
        if (nxt < total) issue(nxt);
        cp_async_commit();
      }
      const int stage = it % NST;
      const uint32_t aBase = smem_u32(sA + stage * A_STAGE);
      const uint32_t a2Base = smem_u32(sA2 + stage * A_STAGE);
      const uint32_t bBase = smem_u32(sB + stage * B_STAGE);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // prologue parameters of the 4 k-columns this thread's A registers cover: k, k+1, k+8, k+9
        float q0[4], q1[4], q2[4];
        if (HAS_P) {
          int kq = kt * BK + ks * 16 + 2 * (lane & 3);
          q0[0] = sP[kq]; q0[1] = sP[kq + 1]; q0[2] = sP[kq + 8]; q0[3] = sP[kq + 9];
          q1[0] = sP[Kpad + kq]; q1[1] = sP[Kpad + kq + 1]; q1[2] = sP[Kpad + kq + 8]; q1[3] = sP[Kpad + kq + 9];
          if (AMODE == CVB_A_BNB) {
            q2[0] = sP[2 * Kpad + kq]; q2[1] = sP[2 * Kpad + kq + 1]; q2[2] = sP[2 * Kpad + kq + 8]; q2[3] = sP[2 * Kpad + kq + 9];
          }
        }
        uint32_t af[MI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
          int row = wm0 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          int ch = ks * 2 + (lane >> 4);
          ldmatrix_x4(aBase + swz64(row, ch), af[mi][0], af[mi][1], af[mi][2], af[mi][3]);
          if (AMODE != CVB_A_RAW) {
            uint32_t a2f[4] = {0, 0, 0, 0};
            if (TWO_A) ldmatrix_x4(a2Base + swz64(row, ch), a2f[0], a2f[1], a2f[2], a2f[3]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              // r: 0 (row g, k lo) 1 (row g+8, k lo) 2 (row g, k hi) 3 (row g+8, k hi)
              const int kk = (r >> 1) * 2;  // index into q*: lo pair -> 0,1 ; hi pair -> 2,3
              const int h = r & 1;          // row half
              float2 x = unpack_bf162(af[mi][r]);
              float y0, y1;
              if (AMODE == CVB_A_AFF) {
                y0 = fmaf(q0[kk], x.x, q1[kk]); y1 = fmaf(q0[kk + 1], x.y, q1[kk + 1]);
              } else if (AMODE == CVB_A_AFF_SILU) {
                y0 = silu_f(fmaf(q0[kk], x.x, q1[kk])); y1 = silu_f(fmaf(q0[kk + 1], x.y, q1[kk + 1]));
              } else if (AMODE == CVB_A_SILU) {
                y0 = silu_f(x.x); y1 = silu_f(x.y);
              } else if (AMODE == CVB_A_GN) {
                float xm0 = (x.x - rmean[mi][h]) * rrstd[mi][h], xm1 = (x.y - rmean[mi][h]) * rrstd[mi][h];
                y0 = fmaf(xm0, q0[kk], q1[kk]); y1 = fmaf(xm1, q0[kk + 1], q1[kk + 1]);
              } else {  // BNB
                float2 x2 = unpack_bf162(a2f[r]);
                y0 = fmaf(q0[kk], x.x, fmaf(q1[kk], x2.x, q2[kk]));
                y1 = fmaf(q0[kk + 1], x.y, fmaf(q1[kk + 1], x2.y, q2[kk + 1]));
              }
              af[mi][r] = pack_bf162(y0, y1);
            }
          }
        }
        uint32_t bfr[4][2];
#pragma unroll
        for (int nj = 0; nj < 2; ++nj) {
          int row = wn0 + nj * 16 + (lane & 7) + (lane >> 4) * 8;
          int ch = ks * 2 + ((lane >> 3) & 1);
          ldmatrix_x4(bBase + swz64(row, ch), bfr[nj * 2][0], bfr[nj * 2][1], bfr[nj * 2 + 1][0], bfr[nj * 2 + 1][1]);
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) mma_bf16_16816(acc[mi][ni], af[mi], bfr[ni][0], bfr[ni][1]);
      }
    }

    // ---------------------------------------------------------------- epilogue: two 64-row halves through the fp32 staging tile
    const int first_sample = m0 / rps;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      __syncthreads();  // staging tile free (previous half / previous tile fully consumed)
      if (wm0 / HALF == half) {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            int r = (wm0 % HALF) + mi * 16 + (lane >> 2);
            int c = wn0 + ni * 8 + 2 * (lane & 3);
            *reinterpret_cast<float2*>(&sC[r * LDC_S + c]) = make_float2(acc[mi][ni][0], acc[mi][ni][1]);
            *reinterpret_cast<float2*>(&sC[(r + 8) * LDC_S + c]) = make_float2(acc[mi][ni][2], acc[mi][ni][3]);
          }
      }
      __syncthreads();
      // per-column epilogue vectors are (re)loaded here (L1 hits) instead of living in registers across the main loop
      float bias8[8], ep0[8], ep1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bias8[j] = (col_ok && p.bias) ? __ldg(p.bias + nc + j) : 0.f;
        ep0[j] = (col_ok && p.e_p0) ? __ldg(p.e_p0 + nc + j) : 1.f;
        ep1[j] = (col_ok && p.e_p1) ? __ldg(p.e_p1 + nc + j) : 0.f;
      }
      for (int r = r0; r < HALF; r += ROWS_PER_PASS) {
        const int m = m0 + half * HALF + r;
        const bool valid = col_ok && (m < p.M);
        float ssum = 0.f, ssq = 0.f;
        if (valid) {
          float v[8];
          float4 t0 = *reinterpret_cast<const float4*>(&sC[r * LDC_S + cg * 8]);
          float4 t1 = *reinterpret_cast<const float4*>(&sC[r * LDC_S + cg * 8 + 4]);
          v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += bias8[j];
          float y8[8];
          if (emode == CVB_E_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
          } else if (emode == CVB_E_SILU_BWD) {
            unpack8(ldg16(Yg + (size_t)m * p.ldy + nc), y8);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= silu_grad_f(fmaf(ep0[j], y8[j], ep1[j]));
          } else if (emode == CVB_E_GN_BWD) {
            unpack8(ldg16(Yg + (size_t)m * p.ldy + nc), y8);
            const int b = m / rps;
            const float mu = p.row_mean[b], rs = p.row_rstd[b];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              y8[j] = (y8[j] - mu) * rs;  // x-hat
              cs[j] += v[j];
              cq[j] += v[j] * y8[j];
              v[j] *= ep0[j];
            }
          }
          if (Rg) {
            float r8[8];
            unpack8(ldg16(Rg + (size_t)m * p.ldr + nc), r8);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += r8[j];
          }
          if (p.c_fp32) {
            float* Cg = static_cast<float*>(p.C) + (size_t)m * p.ldc + nc;
            *reinterpret_cast<float4*>(Cg) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(Cg + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j]);
            stg16(static_cast<bf16*>(p.C) + (size_t)m * p.ldc + nc, pack8(v));
          }
          if (emode == CVB_E_STORE || emode == CVB_E_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { cs[j] += v[j]; cq[j] += v[j] * v[j]; ssum += v[j]; ssq += v[j] * v[j]; }
          } else if (emode == CVB_E_SILU_BWD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { cs[j] += v[j]; cq[j] += v[j] * y8[j]; }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ssum += v[j]; ssq += v[j] * y8[j]; }
          }
        }
        if (want_samp) {
#pragma unroll
          for (int o = CGS / 2; o > 0; o >>= 1) {
            ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
            ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
          }
          if (cg == 0 && m < p.M) {
            int bi = m / rps - first_sample;
            atomicAdd(&s_samp[0][bi], (double)ssum);
            atomicAdd(&s_samp[1][bi], (double)ssq);
          }
        }
      }
    }
    if (want_samp) {
      __syncthreads();  // s_samp complete for this tile
      if (tid < 128) {
        int mlast = min(m0 + BM, p.M) - 1;
        int nsamp = mlast / rps - first_sample + 1;
        if (tid < nsamp) {
          atomicAdd(p.samp_sum + first_sample + tid, s_samp[0][tid]);
          atomicAdd(p.samp_sq + first_sample + tid, s_samp[1][tid]);
          s_samp[0][tid] = 0.0;
          s_samp[1][tid] = 0.0;
        }
      }
    }
  }  // tile loop
  cp_async_wait<0>();

  if (want_col) {
    // reduce over the lanes that share a column group (lane stride CGS), then one smem atomic per warp and column
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = cs[j], q = cq[j];
#pragma unroll
      for (int o = CGS; o < 32; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane < CGS) {
        atomicAdd(&s_col[0][cg * 8 + j], a);
        atomicAdd(&s_col[1][cg * 8 + j], q);
      }
    }
    __syncthreads();
    if (tid < BN && n0 + tid < p.N) {
      atomicAdd(p.col_sum + n0 + tid, (double)s_col[0][tid]);
      atomicAdd(p.col_sq + n0 + tid, (double)s_col[1][tid]);
    }
  }
}

template <int WM, int AMODE>
int launch_gemm(const cvb_gemm_args& a, cudaStream_t st) {
  constexpr int WARPS_M = BM / WM;
  constexpr int BN = (8 / WARPS_M) * 32;
  constexpr int NST = GemmCfg<AMODE>::kStages;
  const int KT = (a.K + BK - 1) / BK;
  const int nvec = (AMODE == CVB_A_AFF || AMODE == CVB_A_AFF_SILU || AMODE == CVB_A_GN) ? 2 : (AMODE == CVB_A_BNB ? 3 : 0);
  size_t smem = (size_t)NST * (BM * BK * 2 * (AMODE == CVB_A_BNB ? 2 : 1) + BN * BK * 2) + (size_t)(BM / 2) * (BN + 4) * 4 +
                (size_t)nvec * KT * BK * 4;
  static bool attr_set = false;
  if (!attr_set) {
    CVB_CUDA(cudaFuncSetAttribute(pw_gemm_kernel<WM, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  CVB_CHECK(smem <= 200 * 1024, "cvb_pw_gemm: K=%d too large for the prologue parameter cache", a.K);
  int occ = 0;
  CVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pw_gemm_kernel<WM, AMODE>, NTHREADS, smem));
  if (occ < 1) occ = 1;
  const int n_tiles = (a.N + BN - 1) / BN, m_tiles = (a.M + BM - 1) / BM;
  int gy = (occ * cvb_num_sms() + n_tiles - 1) / n_tiles;  // all CTAs resident, each streaming over its M tiles
  if (gy > m_tiles) gy = m_tiles;
  if (gy < 1) gy = 1;
  dim3 grid(n_tiles, gy);
  pw_gemm_kernel<WM, AMODE><<<grid, NTHREADS, smem, st>>>(a);
  CVB_LAUNCH_CHECK();
  return 0;
}

template <int AMODE>
int dispatch_tile(const cvb_gemm_args& a, cudaStream_t st) {
  int N = a.N;
  if (N <= 32) return launch_gemm<16, AMODE>(a, st);
  int pad128 = (N + 127) / 128 * 128, pad64 = (N + 63) / 64 * 64;
  if (N <= 64 || pad64 < pad128) return launch_gemm<32, AMODE>(a, st);
  return launch_gemm<64, AMODE>(a, st);
}

// =====================================================================================================================
// weight gradient
// =====================================================================================================================
constexpr int WG_TN = 64, WG_TK = 64, WG_MB = 64, WG_STAGES = 4;

__device__ __forceinline__ uint32_t swz128(int row, int ch) {  // 128-byte rows, 8 x 16B chunks
  return static_cast<uint32_t>(row * 128 + ((ch ^ (row & 7)) << 4));
}

template <int GMODE, int AMODE>
__global__ void __launch_bounds__(NTHREADS) pw_wgrad_kernel(const cvb_wgrad_args p, int m_per_cta) {
  constexpr bool TWO_G = (GMODE == CVB_A_BNB);
  constexpr int T_STAGE = WG_MB * 64 * 2;  // bytes per operand tile and stage
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sG = smem;
  uint8_t* sG2 = smem + WG_STAGES * T_STAGE;
  uint8_t* sA = smem + (TWO_G ? 2 : 1) * WG_STAGES * T_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wn0 = (warp >> 2) * 32;  // 2 warps along n
  const int wk0 = (warp & 3) * 16;   // 4 warps along k
  const int k0 = blockIdx.x * WG_TK;
  const int n0 = blockIdx.y * WG_TN;
  const int m_begin = blockIdx.z * m_per_cta;
  const int m_end = min(p.M, m_begin + m_per_cta);
  const int NS = (m_end - m_begin + WG_MB - 1) / WG_MB;
  if (NS <= 0) return;

  const bf16* __restrict__ G = static_cast<const bf16*>(p.G);
  const bf16* __restrict__ G2 = static_cast<const bf16*>(p.G2);
  const bf16* __restrict__ A = static_cast<const bf16*>(p.A);

  auto load_stage = [&](int s, int stage) {
#pragma unroll
    for (int i = 0; i < WG_MB / 32; ++i) {
      const int row = (tid >> 3) + i * 32, ch = tid & 7;
      const int m = m_begin + s * WG_MB + row;
      {
        int n = n0 + ch * 8;
        bool ok = (m < m_end) && (n < p.N);
        cp_async16(smem_u32(sG + stage * T_STAGE) + swz128(row, ch), G + (ok ? (size_t)m * p.ldg + n : 0), ok);
        if (TWO_G) cp_async16(smem_u32(sG2 + stage * T_STAGE) + swz128(row, ch), G2 + (ok ? (size_t)m * p.ldg2 + n : 0), ok);
      }
      {
        int k = k0 + ch * 8;
        bool ok = (m < m_end) && (k < p.K);
        cp_async16(smem_u32(sA + stage * T_STAGE) + swz128(row, ch), A + (ok ? (size_t)m * p.lda + k : 0), ok);
      }
    }
  };

  // per-n parameters (rows g, g+8 of the two m16 tiles) and per-k parameters (cols g of the two n8 tiles)
  float gp0[2][2], gp1[2][2], gp2[2][2];
#pragma unroll
  for (int ni = 0; ni < 2; ++ni)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + wn0 + ni * 16 + g + h * 8;
      bool ok = TWO_G && n < p.N;
      gp0[ni][h] = ok ? p.g_p0[n] : 0.f;
      gp1[ni][h] = ok ? p.g_p1[n] : 0.f;
      gp2[ni][h] = ok ? p.g_p2[n] : 0.f;
    }
  float ap0[2], ap1[2];
#pragma unroll
  for (int kj = 0; kj < 2; ++kj) {
    int k = k0 + wk0 + kj * 8 + g;
    bool ok = (AMODE == CVB_A_AFF || AMODE == CVB_A_AFF_SILU || AMODE == CVB_A_GN) && k < p.K;
    ap0[kj] = ok ? p.a_p0[k] : 0.f;
    ap1[kj] = ok ? p.a_p1[k] : 0.f;
  }
  const int rps = p.rows_per_sample > 0 ? p.rows_per_sample : 1;
  const bool want_db = (p.dbias != nullptr) && (blockIdx.x == 0) && ((warp & 3) == 0);
  float db[2][2] = {{0.f, 0.f}, {0.f, 0.f}};

  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

#pragma unroll
  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < NS) load_stage(s, s);
    cp_async_commit();
  }
  for (int s = 0; s < NS; ++s) {
    cp_async_wait<WG_STAGES - 2>();
    __syncthreads();
    {
      int ns = s + WG_STAGES - 1;
      if (ns < NS) load_stage(ns, ns % WG_STAGES);
      cp_async_commit();
    }
    const int stage = s % WG_STAGES;
    const uint32_t gBase = smem_u32(sG + stage * T_STAGE), g2Base = smem_u32(sG2 + stage * T_STAGE);
    const uint32_t aBase = smem_u32(sA + stage * T_STAGE);
    const int ms0 = m_begin + s * WG_MB;
    const bool tail = (ms0 + WG_MB > m_end);
#pragma unroll
    for (int ms = 0; ms < WG_MB / 16; ++ms) {
      // G' fragments (mma A operand: rows = n, cols = m), two m16 tiles
      uint32_t gf[2][4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        int row = ms * 16 + (lane & 7) + (lane >> 4) * 8;
        int ch = (wn0 + ni * 16) / 8 + ((lane >> 3) & 1);
        ldmatrix_x4_trans(gBase + swz128(row, ch), gf[ni][0], gf[ni][1], gf[ni][2], gf[ni][3]);
        if (TWO_G || tail || want_db) {
          uint32_t g2f[4] = {0, 0, 0, 0};
          if (TWO_G) ldmatrix_x4_trans(g2Base + swz128(row, ch), g2f[0], g2f[1], g2f[2], g2f[3]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            // r: 0 (n g, m lo) 1 (n g+8, m lo) 2 (n g, m hi) 3 (n g+8, m hi)
            const int h = r & 1;
            float2 x = unpack_bf162(gf[ni][r]);
            if (TWO_G) {
              float2 x2 = unpack_bf162(g2f[r]);
              x.x = fmaf(gp0[ni][h], x.x, fmaf(gp1[ni][h], x2.x, gp2[ni][h]));
              x.y = fmaf(gp0[ni][h], x.y, fmaf(gp1[ni][h], x2.y, gp2[ni][h]));
            }
            if (tail) {
              int mm = ms0 + ms * 16 + (r >> 1) * 8 + 2 * t;
              if (mm >= m_end) x.x = 0.f;
              if (mm + 1 >= m_end) x.y = 0.f;
            }
            uint32_t pk = pack_bf162(x.x, x.y);
            gf[ni][r] = pk;
            if (want_db) {
              float2 xr = unpack_bf162(pk);
              db[ni][h] += xr.x + xr.y;
            }
          }
        }
      }
      // A' fragments (mma B operand: k16 = m, n8 = k), two n8 tiles from one x4
      uint32_t af[4];
      {
        int row = ms * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        int ch = wk0 / 8 + (lane >> 4);
        ldmatrix_x4_trans(aBase + swz128(row, ch), af[0], af[1], af[2], af[3]);
        if (AMODE != CVB_A_RAW) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            // r: 0 (kj0, m lo) 1 (kj0, m hi) 2 (kj1, m lo) 3 (kj1, m hi); element pair = m 2t, 2t+1
            const int kj = r >> 1;
            float2 x = unpack_bf162(af[r]);
            if (AMODE == CVB_A_GN) {
              int mm = ms0 + ms * 16 + (r & 1) * 8 + 2 * t;
              int b0 = min(mm, p.M - 1) / rps, b1 = min(mm + 1, p.M - 1) / rps;
              x.x = fmaf((x.x - p.row_mean[b0]) * p.row_rstd[b0], ap0[kj], ap1[kj]);
              x.y = fmaf((x.y - p.row_mean[b1]) * p.row_rstd[b1], ap0[kj], ap1[kj]);
            } else {
              x.x = apply_mode(AMODE, x.x, ap0[kj], ap1[kj]);
              x.y = apply_mode(AMODE, x.y, ap0[kj], ap1[kj]);
            }
            af[r] = pack_bf162(x.x, x.y);
          }
        }
      }
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
#pragma unroll
        for (int kj = 0; kj < 2; ++kj) mma_bf16_16816(acc[ni][kj], gf[ni], af[kj * 2], af[kj * 2 + 1]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int ni = 0; ni < 2; ++ni)
#pragma unroll
    for (int kj = 0; kj < 2; ++kj)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int n = n0 + wn0 + ni * 16 + g + (e >> 1) * 8;
        int k = k0 + wk0 + kj * 8 + 2 * t + (e & 1);
        if (n < p.N && k < p.K) atomicAdd(p.dW + (size_t)n * p.lddw + k, acc[ni][kj][e]);
      }
  if (want_db) {
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = db[ni][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        int n = n0 + wn0 + ni * 16 + g + h * 8;
        if (t == 0 && n < p.N) atomicAdd(p.dbias + n, v);
      }
  }
}

template <int GMODE, int AMODE>
int launch_wgrad(const cvb_wgrad_args& a, cudaStream_t st) {
  const int kt = (a.K + WG_TK - 1) / WG_TK, nt = (a.N + WG_TN - 1) / WG_TN;
  const int target = 4 * cvb_num_sms();
  int splits = (target + kt * nt - 1) / (kt * nt);
  int max_splits = (a.M + 255) / 256;  // at least 256 rows per CTA
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_cta = ((a.M + splits - 1) / splits + WG_MB - 1) / WG_MB * WG_MB;
  splits = (a.M + m_per_cta - 1) / m_per_cta;
  size_t smem = (size_t)WG_STAGES * WG_MB * 64 * 2 * (GMODE == CVB_A_BNB ? 3 : 2);
  static bool attr_set = false;
  if (!attr_set) {
    CVB_CUDA(cudaFuncSetAttribute(pw_wgrad_kernel<GMODE, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_set = true;
  }
  dim3 grid(kt, nt, splits);
  pw_wgrad_kernel<GMODE, AMODE><<<grid, NTHREADS, smem, st>>>(a, m_per_cta);
  CVB_LAUNCH_CHECK();
  return 0;
}

template <int GMODE>
int dispatch_wgrad_a(const cvb_wgrad_args& a, cudaStream_t st) {
  switch (a.a_mode) {
    case CVB_A_RAW: return launch_wgrad<GMODE, CVB_A_RAW>(a, st);
    case CVB_A_AFF: return launch_wgrad<GMODE, CVB_A_AFF>(a, st);
    case CVB_A_AFF_SILU: return launch_wgrad<GMODE, CVB_A_AFF_SILU>(a, st);
    case CVB_A_SILU: return launch_wgrad<GMODE, CVB_A_SILU>(a, st);
    case CVB_A_GN: return launch_wgrad<GMODE, CVB_A_GN>(a, st);
    default: cvb_set_error("cvb_pw_wgrad: unsupported a_mode %d", a.a_mode); return 1;
  }
}

}  // namespace

extern "C" int cvb_pw_gemm(const cvb_gemm_args* args, cvb_stream_t stream) {
  CVB_CHECK(args != nullptr, "cvb_pw_gemm: null args");
  const cvb_gemm_args& a = *args;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CVB_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "cvb_pw_gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  CVB_CHECK(a.K % 8 == 0 && a.N % 8 == 0, "cvb_pw_gemm: K (%d) and N (%d) must be multiples of 8", a.K, a.N);
  CVB_CHECK(a.lda % 8 == 0 && a.ldw % 8 == 0 && a.ldc % (a.c_fp32 ? 4 : 8) == 0, "cvb_pw_gemm: leading dims must be multiples of 8");
  CVB_CHECK(a.A && a.W && a.C, "cvb_pw_gemm: null operand");
  CVB_CHECK(cvb_aligned16(a.A) && cvb_aligned16(a.W) && cvb_aligned16(a.C), "cvb_pw_gemm: operands must be 16-byte aligned");
  CVB_CHECK(a.e_mode >= CVB_E_STORE && a.e_mode <= CVB_E_GN_BWD, "cvb_pw_gemm: bad e_mode %d", a.e_mode);
  if (a.e_mode == CVB_E_SILU_BWD || a.e_mode == CVB_E_GN_BWD)
    CVB_CHECK(a.Y && a.ldy % 8 == 0 && cvb_aligned16(a.Y), "cvb_pw_gemm: epilogue mode %d needs Y", a.e_mode);
  if (a.e_mode == CVB_E_GN_BWD || a.a_mode == CVB_A_GN)
    CVB_CHECK(a.row_mean && a.row_rstd && a.rows_per_sample > 0, "cvb_pw_gemm: GroupNorm modes need row_mean/row_rstd/rows_per_sample");
  if (a.R) CVB_CHECK(a.ldr % 8 == 0 && cvb_aligned16(a.R), "cvb_pw_gemm: bad residual");
  if (a.samp_sum) CVB_CHECK(a.samp_sq && a.rows_per_sample > 0, "cvb_pw_gemm: sample statistics need rows_per_sample");
  if (a.col_sum) CVB_CHECK(a.col_sq != nullptr, "cvb_pw_gemm: col_sq missing");
  switch (a.a_mode) {
    case CVB_A_RAW: return dispatch_tile<CVB_A_RAW>(a, st);
    case CVB_A_AFF: CVB_CHECK(a.a_p0 && a.a_p1, "cvb_pw_gemm: AFF needs p0/p1"); return dispatch_tile<CVB_A_AFF>(a, st);
    case CVB_A_AFF_SILU: CVB_CHECK(a.a_p0 && a.a_p1, "cvb_pw_gemm: AFF_SILU needs p0/p1"); return dispatch_tile<CVB_A_AFF_SILU>(a, st);
    case CVB_A_SILU: return dispatch_tile<CVB_A_SILU>(a, st);
    case CVB_A_GN: CVB_CHECK(a.a_p0 && a.a_p1, "cvb_pw_gemm: GN needs gamma/beta"); return dispatch_tile<CVB_A_GN>(a, st);
    case CVB_A_BNB:
      CVB_CHECK(a.A2 && a.a_p0 && a.a_p1 && a.a_p2 && a.lda2 % 8 == 0 && cvb_aligned16(a.A2), "cvb_pw_gemm: BNB needs A2 and p0/p1/p2");
      return dispatch_tile<CVB_A_BNB>(a, st);
    default: cvb_set_error("cvb_pw_gemm: unsupported a_mode %d", a.a_mode); return 1;
  }
}

extern "C" int cvb_pw_wgrad(const cvb_wgrad_args* args, cvb_stream_t stream) {
  CVB_CHECK(args != nullptr, "cvb_pw_wgrad: null args");
  const cvb_wgrad_args& a = *args;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CVB_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "cvb_pw_wgrad: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  CVB_CHECK(a.K % 8 == 0 && a.N % 8 == 0, "cvb_pw_wgrad: K (%d) and N (%d) must be multiples of 8", a.K, a.N);
  CVB_CHECK(a.G && a.A && a.dW, "cvb_pw_wgrad: null operand");
  CVB_CHECK(a.ldg % 8 == 0 && a.lda % 8 == 0 && cvb_aligned16(a.G) && cvb_aligned16(a.A), "cvb_pw_wgrad: operands must be 16-byte aligned / ld % 8");
  if (a.a_mode == CVB_A_GN) CVB_CHECK(a.row_mean && a.row_rstd && a.rows_per_sample > 0 && a.a_p0 && a.a_p1, "cvb_pw_wgrad: GN needs statistics");
  if (a.a_mode == CVB_A_AFF || a.a_mode == CVB_A_AFF_SILU) CVB_CHECK(a.a_p0 && a.a_p1, "cvb_pw_wgrad: AFF needs p0/p1");
  if (a.g_mode == CVB_A_RAW) return dispatch_wgrad_a<CVB_A_RAW>(a, st);
  if (a.g_mode == CVB_A_BNB) {
    CVB_CHECK(a.G2 && a.g_p0 && a.g_p1 && a.g_p2 && a.ldg2 % 8 == 0 && cvb_aligned16(a.G2), "cvb_pw_wgrad: BNB needs G2 and p0/p1/p2");
    return dispatch_wgrad_a<CVB_A_BNB>(a, st);
  }
  cvb_set_error("cvb_pw_wgrad: unsupported g_mode %d", a.g_mode);
  return 1;
}
