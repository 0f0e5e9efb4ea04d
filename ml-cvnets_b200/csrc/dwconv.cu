// Depthwise 3x3 convolution (pad 1, stride 1|2), NHWC bf16, forward and fused backward (sm_100a).
//
// Pure HBM-bound stencils (9 MAC per element): the kernels stage a spatial tile x 64 channels (+halo) in shared memory
// with the PRODUCER's BatchNorm(+SiLU) already applied (so every input element is transformed exactly once), use
// 16-byte channel vectors everywhere, and emit the BatchNorm statistics of their own output in the epilogue.
// The backward kernel fuses  dy = BN-backward(dz, y)  ->  dX (transposed stencil)  ->  activation backward of the
// producer + its BN-backward statistics  and  dW (per-channel 9-tap reduction) into one pass over the tensors.
#include "common.cuh"

namespace {

constexpr int CB = 64;  // channels per CTA (8 x 16-byte chunks)
constexpr int NT = 256;

// smem tiles are plain [pixel][64 channels] (128 B per pixel, no swizzle): every access pattern below touches one pixel's 128 B per
// quarter-warp / warp and is conflict-free as is, and linear addresses keep the address arithmetic out of the instruction stream
__device__ __forceinline__ uint32_t pix_off(int pix, int ch) { return static_cast<uint32_t>(pix * 128 + (ch << 4)); }

__device__ __forceinline__ void load8_mode(int mode, const bf16* ptr, const float* p0, const float* p1, float* out) {
  unpack8(ldg16(ptr), out);
  if (mode == CVB_A_AFF) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = fmaf(p0[j], out[j], p1[j]);
  } else if (mode == CVB_A_AFF_SILU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = silu_f(fmaf(p0[j], out[j], p1[j]));
  }
}

// ------------------------------------------------------------------------------------------------------------- forward
// TMA-staged: one elected thread issues a 4-D tensor load of the [IH, IW, 64ch] halo tile (out-of-bounds = zero fill = the
// conv padding) straight into 128B-swizzled shared memory; two buffers + mbarriers keep the NEXT image's tile in flight while
// the current one is transformed (producer BN+SiLU, in place) and convolved.
template <int XMODE>
__global__ void __launch_bounds__(NT, 2) dw_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const cvb_dw_fwd_args p, int Ho, int Wo, int TH,
                                                       int TW, int logTW, int tiles_w, int buf_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B-swizzled TMA destinations must be 1024-byte aligned in the shared window: align by hand (host adds 1 KB of slack)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ float s_cs[CB], s_cq[CB];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = p.stride;
  const int th_i = blockIdx.x / tiles_w, tw_i = blockIdx.x % tiles_w;
  const int oh0 = th_i * TH, ow0 = tw_i * TW;
  const int c0 = blockIdx.y * CB;
  const int IH = (TH - 1) * s + 3, IW = (TW - 1) * s + 3;
  const int h_base = oh0 * s - 1, w_base = ow0 * s - 1;
  const uint32_t tile_bytes = (uint32_t)IH * IW * 128;
  const int n_img = (p.B - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;

  if (tid < CB) { s_cs[tid] = 0.f; s_cq[tid] = 0.f; }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < 2 && i < n_img; ++i) {
      mbar_expect_tx(&bar[i], tile_bytes);
      tma_load_4d(smem + i * buf_bytes, &tmX, &bar[i], c0, w_base, h_base, (int)blockIdx.z + i * (int)gridDim.z);
    }
  }

  const int cgi = tid & 7, pt = tid >> 3;
  const int c = c0 + cgi * 8;
  const bool c_ok = c < p.C;
  float cs[8], cq[8];  // BatchNorm statistics, accumulated over the batch loop and flushed once per CTA
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
  uint32_t wpk[9][4];  // the 72 weights of this thread's 8 channels are bf16 values: keep them packed
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int j = 0; j < 4; ++j) wpk[tp][j] = c_ok ? pack_bf162(p.Wt[tp * p.C + c + 2 * j], p.Wt[tp * p.C + c + 2 * j + 1]) : 0u;

  // producer BN scale/shift of this thread's 8 channels (transform role: chunk = lane & 7 == cgi), hoisted out of all loops
  float xp0[8], xp1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xp0[j] = (XMODE != CVB_A_RAW && c_ok) ? __ldg(p.x_p0 + c + j) : 1.f;
    xp1[j] = (XMODE != CVB_A_RAW && c_ok) ? __ldg(p.x_p1 + c + j) : 0.f;
  }
  for (int i = 0; i < n_img; ++i) {
    const int b = (int)blockIdx.z + i * (int)gridDim.z;
    uint8_t* tile = smem + (i & 1) * buf_bytes;
    mbar_wait(&bar[i & 1], (i >> 1) & 1);
    if (XMODE != CVB_A_RAW) {
      // in-place producer transform; out-of-bounds pixels / channels stay zero (padding acts on the activated tensor)
      const int pch = lane & 7;  // physical 16-byte chunk inside the pixel's 128-byte row
      for (int ih = warp; ih < IH; ih += NT / 32) {
        const int h = h_base + ih;
        if (h < 0 || h >= p.H) continue;
        for (int jw = lane >> 3; jw < IW; jw += 4) {
          const int w = w_base + jw;
          const int pix = ih * IW + jw;
          const int lc = c0 + (pch << 3);
          if (w < 0 || w >= p.W || lc >= p.C) continue;
          uint4* ptr = reinterpret_cast<uint4*>(tile + pix * 128 + (pch << 4));
          float f[8];
          unpack8(*ptr, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float z = fmaf(xp0[j], f[j], xp1[j]);
            f[j] = (XMODE == CVB_A_AFF_SILU) ? silu_f(z) : z;
          }
          *ptr = pack8(f);
        }
      }
      __syncthreads();
    }
    bf16* __restrict__ Y = static_cast<bf16*>(p.Y) + (size_t)b * Ho * Wo * p.C;
    for (int op = pt; op < TH * TW; op += NT / 8) {
      const int oh = op >> logTW, ow = op & (TW - 1);
      const int gh = oh0 + oh, gw = ow0 + ow;
      if (gh < Ho && gw < Wo && c_ok) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            float xin[8];
            unpack8(*reinterpret_cast<const uint4*>(tile + pix_off((oh * s + u) * IW + ow * s + v, cgi)), xin);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 wv = unpack_bf162(wpk[u * 3 + v][j]);
              acc[2 * j] = fmaf(wv.x, xin[2 * j], acc[2 * j]);
              acc[2 * j + 1] = fmaf(wv.y, xin[2 * j + 1], acc[2 * j + 1]);
            }
          }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = bf16_round(acc[j]);
          cs[j] += acc[j];
          cq[j] += acc[j] * acc[j];
        }
        stg16(Y + ((size_t)gh * Wo + gw) * p.C + c, pack8(acc));
      }
    }
    __syncthreads();  // every thread is done with this buffer
    if (tid == 0 && i + 2 < n_img) {
      fence_proxy_async();  // order the generic-proxy accesses above before the async-proxy overwrite
      mbar_expect_tx(&bar[i & 1], tile_bytes);
      tma_load_4d(tile, &tmX, &bar[i & 1], c0, w_base, h_base, (int)blockIdx.z + (i + 2) * (int)gridDim.z);
    }
  }
  if (p.col_sum) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // reduce over the 4 pixel-threads of this warp that share the channel chunk (lane bits 3,4)
      float a = cs[j], q = cq[j];
      a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
      q += __shfl_xor_sync(0xffffffffu, q, 8); q += __shfl_xor_sync(0xffffffffu, q, 16);
      if (lane < 8) { atomicAdd(&s_cs[cgi * 8 + j], a); atomicAdd(&s_cq[cgi * 8 + j], q); }
    }
    __syncthreads();
    if (tid < CB && c0 + tid < p.C) {
      atomicAdd(p.col_sum + c0 + tid, (double)s_cs[tid]);
      atomicAdd(p.col_sq + c0 + tid, (double)s_cq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------ backward
// One pass over (dz, y2, x) per image tile, all three staged by TMA (zero-filled halos) into two buffer sets so the next
// image's tiles are in flight while this one is processed:
//   1. dy = BN-backward(dz, y2) in place                      (generic transform of the dz tile)
//   2. phase B: dX = conv_transpose(dy) * silu'(BN(x)) (+ the producer's BN-backward statistics), raw x read from the smem tile
//   3. a = SiLU(BN(x)) in place, phase A: dW[9 taps] += dy * a(shifted)
// 512 threads, register-lean roles: phase A thread = (channel pair, pixel slice = warp) with 18 accumulators kept over the batch
// loop; phase B thread = (8-channel chunk, pixel) with the 72 weights as 36 packed bf16x2 registers.
constexpr int NTB = 512;

template <int GMODE, int XMODE>
__global__ void __launch_bounds__(NTB, 1) dw_bwd_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmY2,
                                                        const __grid_constant__ CUtensorMap tmX, const cvb_dw_bwd_args p, int Ho, int Wo, int TH,
                                                        int TW, int logTW, int tiles_w, int g_bytes, int x_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ float s_cs[CB], s_cq[CB];
  __shared__ float s_dw[9][CB];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = p.stride;
  const int th_i = blockIdx.x / tiles_w, tw_i = blockIdx.x % tiles_w;
  const int oh0 = th_i * TH, ow0 = tw_i * TW;
  const int c0 = blockIdx.y * CB;
  const int go = (s == 1) ? 1 : 0;          // halo of the dy tile on the low side
  const int GH = TH + (s == 1 ? 2 : 1), GW = TW + (s == 1 ? 2 : 1);
  const int XH = s * TH + 3 - s, XW = s * TW + 3 - s;  // input tile + halo (origin -1)
  const int ITH = s * TH, ITW = s * TW;     // owned input tile
  const int logITW = logTW + (s == 2 ? 1 : 0);
  const int set_bytes = (GMODE == CVB_A_BNB ? 2 : 1) * g_bytes + x_bytes;
  const uint32_t tx_bytes = (uint32_t)(GMODE == CVB_A_BNB ? 2 : 1) * GH * GW * 128 + (uint32_t)XH * XW * 128;
  const int n_img = (p.B - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;
  const int gh_base = oh0 - go, gw_base = ow0 - go;
  const int xh_base = s * oh0 - 1, xw_base = s * ow0 - 1;

  auto issue = [&](int i) {  // one elected thread
    uint8_t* base = smem + (i & 1) * set_bytes;
    const int b = (int)blockIdx.z + i * (int)gridDim.z;
    mbar_expect_tx(&bar[i & 1], tx_bytes);
    tma_load_4d(base, &tmDZ, &bar[i & 1], c0, gw_base, gh_base, b);
    if (GMODE == CVB_A_BNB) tma_load_4d(base + g_bytes, &tmY2, &bar[i & 1], c0, gw_base, gh_base, b);
    tma_load_4d(base + (GMODE == CVB_A_BNB ? 2 : 1) * g_bytes, &tmX, &bar[i & 1], c0, xw_base, xh_base, b);
  };

  for (int i = tid; i < 9 * CB; i += NTB) (&s_dw[0][0])[i] = 0.f;
  if (tid < CB) { s_cs[tid] = 0.f; s_cq[tid] = 0.f; }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < 2 && i < n_img; ++i) issue(i);
  }

  const int cgi = tid & 7, pt = tid >> 3;   // phase B role
  const int cc = c0 + cgi * 8;
  const bool cc_ok = cc < p.C;
  const int cp = lane, ps = warp;           // phase A role: channels c0 + 2cp, +1 ; pixel slice = warp
  const int pch = lane & 7;                 // transform role: physical 16-byte chunk

  float accw[9][2];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) { accw[tp][0] = 0.f; accw[tp][1] = 0.f; }
  float cs[8], cq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
  uint32_t wpk[9][4];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      wpk[tp][j] = cc_ok ? pack_bf162(p.Wt[tp * p.C + cc + 2 * j], p.Wt[tp * p.C + cc + 2 * j + 1]) : 0u;

  // producer BN scale/shift of the CTA's 64 channels live in smem (the register budget of 512 threads is spent on the stencil)
  __shared__ __align__(16) float s_xp[2][CB];
  if (tid < CB) {
    const bool ok = (XMODE != CVB_A_RAW) && (c0 + tid < p.C);
    s_xp[0][tid] = ok ? __ldg(p.x_p0 + c0 + tid) : 1.f;
    s_xp[1][tid] = ok ? __ldg(p.x_p1 + c0 + tid) : 0.f;
  }
  __syncthreads();
  for (int i = 0; i < n_img; ++i) {
    const int b = (int)blockIdx.z + i * (int)gridDim.z;
    uint8_t* sG = smem + (i & 1) * set_bytes;
    uint8_t* sG2 = sG + g_bytes;
    uint8_t* sX = sG + (GMODE == CVB_A_BNB ? 2 : 1) * g_bytes;
    bf16* __restrict__ DX = static_cast<bf16*>(p.DX) + (size_t)b * p.H * p.W * p.C;
    mbar_wait(&bar[i & 1], (i >> 1) & 1);
    // ---- 1. dy = c1*dz + c2*y2 + c3, in place, in-bounds pixels only (the zero-filled halo must stay zero)
    if (GMODE == CVB_A_BNB) {
      float g0[8], g1[8], g2[8];  // BN-backward coefficients of this thread's chunk (scoped: live only during this pass)
      const int glc = c0 + (pch << 3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = glc < p.C;
        g0[j] = ok ? __ldg(p.g_p0 + glc + j) : 0.f;
        g1[j] = ok ? __ldg(p.g_p1 + glc + j) : 0.f;
        g2[j] = ok ? __ldg(p.g_p2 + glc + j) : 0.f;
      }
      for (int gi = warp; gi < GH; gi += NTB / 32) {
        const int oh = gh_base + gi;
        if (oh < 0 || oh >= Ho) continue;
        for (int gj = lane >> 3; gj < GW; gj += 4) {
          const int ow = gw_base + gj;
          const int pix = gi * GW + gj;
          const int lc = c0 + (pch << 3);
          if (ow < 0 || ow >= Wo || lc >= p.C) continue;
          uint4* pz = reinterpret_cast<uint4*>(sG + pix * 128 + (pch << 4));
          float f[8], y[8];
          unpack8(*pz, f);
          unpack8(*reinterpret_cast<const uint4*>(sG2 + pix * 128 + (pch << 4)), y);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(g0[j], f[j], fmaf(g1[j], y[j], g2[j]));
          *pz = pack8(f);
        }
      }
      __syncthreads();
    }
    // ---- 2. phase B: input gradient  da[h,w] = sum_{u,v} W[u,v] * dy[(h+1-u)/s, (w+1-v)/s]
    for (int ip = pt; ip < ITH * ITW; ip += NTB / 8) {
      const int ih = ip >> logITW, iw = ip & (ITW - 1);
      const int h = s * oh0 + ih, w = s * ow0 + iw;
      if (h < p.H && w < p.W && cc_ok) {
        float da[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) da[j] = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          int gi;
          if (s == 1) gi = ih + 2 - u;
          else { if (((ih + 1 - u) & 1) != 0) continue; gi = (ih + 1 - u) >> 1; }
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            int gj;
            if (s == 1) gj = iw + 2 - v;
            else { if (((iw + 1 - v) & 1) != 0) continue; gj = (iw + 1 - v) >> 1; }
            float dy[8];
            unpack8(*reinterpret_cast<const uint4*>(sG + pix_off(gi * GW + gj, cgi)), dy);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 wv = unpack_bf162(wpk[u * 3 + v][j]);
              da[2 * j] = fmaf(wv.x, dy[2 * j], da[2 * j]);
              da[2 * j + 1] = fmaf(wv.y, dy[2 * j + 1], da[2 * j + 1]);
            }
          }
        }
        if (XMODE != CVB_A_RAW) {
          // one sigmoid per element serves both uses: a = z*s (kept in place for phase A) and silu'(z) = s + a*(1-s)
          uint4* px = reinterpret_cast<uint4*>(sX + pix_off((ih + 1) * XW + iw + 1, cgi));
          float xr[8], av[8], xp0[8], xp1[8];
          unpack8(*px, xr);
          *reinterpret_cast<float4*>(xp0) = *reinterpret_cast<const float4*>(&s_xp[0][cgi * 8]);
          *reinterpret_cast<float4*>(xp0 + 4) = *reinterpret_cast<const float4*>(&s_xp[0][cgi * 8 + 4]);
          *reinterpret_cast<float4*>(xp1) = *reinterpret_cast<const float4*>(&s_xp[1][cgi * 8]);
          *reinterpret_cast<float4*>(xp1 + 4) = *reinterpret_cast<const float4*>(&s_xp[1][cgi * 8 + 4]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = fmaf(xp0[j], xr[j], xp1[j]);
            if (XMODE == CVB_A_AFF_SILU) {
              const float sg = 1.0f / (1.0f + __expf(-z));
              av[j] = z * sg;
              da[j] *= fmaf(av[j], 1.0f - sg, sg);
            } else {
              av[j] = z;
            }
            da[j] = bf16_round(da[j]);
            cs[j] += da[j];
            cq[j] += da[j] * xr[j];
          }
          *px = pack8(av);
        }
        stg16(DX + ((size_t)h * p.W + w) * p.C + cc, pack8(da));
      }
    }
    // ---- 3. halo ring of the input tile: a = act(BN(x)) in place (in-bounds only; the centre was done in phase B), then phase A
    if (XMODE != CVB_A_RAW) {
      for (int ih = warp; ih < XH; ih += NTB / 32) {
        const int h = xh_base + ih;
        if (h < 0 || h >= p.H) continue;
        const bool row_center = (ih >= 1 && ih <= ITH);
        for (int jw = lane >> 3; jw < XW; jw += 4) {
          if (row_center && jw >= 1 && jw <= ITW) continue;
          const int w = xw_base + jw;
          const int pix = ih * XW + jw;
          const int lc = c0 + (pch << 3);
          if (w < 0 || w >= p.W || lc >= p.C) continue;
          uint4* px = reinterpret_cast<uint4*>(sX + pix * 128 + (pch << 4));
          float f[8];
          unpack8(*px, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float z = fmaf(s_xp[0][pch * 8 + j], f[j], s_xp[1][pch * 8 + j]);
            f[j] = (XMODE == CVB_A_AFF_SILU) ? silu_f(z) : z;
          }
          *px = pack8(f);
        }
      }
    }
    __syncthreads();  // dX done with the dy tile; transformed input tile complete
    {
      const uint32_t sub = static_cast<uint32_t>((cp & 3) << 2);
      const int chk = cp >> 2;
      for (int op = ps; op < TH * TW; op += NTB / 32) {
        const int oh = op >> logTW, ow = op & (TW - 1);
        const float2 dy = unpack_bf162(*reinterpret_cast<const uint32_t*>(sG + pix_off((oh + go) * GW + ow + go, chk) + sub));
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const float2 a = unpack_bf162(*reinterpret_cast<const uint32_t*>(sX + pix_off((s * oh + u) * XW + s * ow + v, chk) + sub));
            accw[u * 3 + v][0] = fmaf(dy.x, a.x, accw[u * 3 + v][0]);
            accw[u * 3 + v][1] = fmaf(dy.y, a.y, accw[u * 3 + v][1]);
          }
      }
    }
    __syncthreads();  // every thread is done with this buffer set
    if (tid == 0 && i + 2 < n_img) {
      fence_proxy_async();
      issue(i + 2);
    }
  }

  // ---- reductions (once per CTA)
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    atomicAdd(&s_dw[tp][2 * cp], accw[tp][0]);
    atomicAdd(&s_dw[tp][2 * cp + 1], accw[tp][1]);
  }
  if (p.col_sum) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = cs[j], q = cq[j];
      a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
      q += __shfl_xor_sync(0xffffffffu, q, 8); q += __shfl_xor_sync(0xffffffffu, q, 16);
      if (lane < 8) { atomicAdd(&s_cs[cgi * 8 + j], a); atomicAdd(&s_cq[cgi * 8 + j], q); }
    }
  }
  __syncthreads();
  for (int i = tid; i < 9 * CB; i += NTB) {
    int tp = i / CB, c = i % CB;
    if (c0 + c < p.C) atomicAdd(p.dWt + tp * p.C + c0 + c, s_dw[tp][c]);
  }
  if (p.col_sum && tid < CB && c0 + tid < p.C) {
    atomicAdd(p.col_sum + c0 + tid, (double)s_cs[tid]);
    atomicAdd(p.col_sq + c0 + tid, (double)s_cq[tid]);
  }
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

void pick_tile(int Ho, int Wo, int stride, int* TH, int* TW) {
  int tw = 1 << ilog2(Wo); if (tw > 16) tw = 16;
  int budget = (stride == 1 ? 128 : 64) / tw;
  int th = 1 << ilog2(Ho); if (th > budget) th = budget; if (th < 1) th = 1;
  *TH = th; *TW = tw;
}

}  // namespace

extern "C" int cvb_dw_fwd(const cvb_dw_fwd_args* args, cvb_stream_t stream) {
  CVB_CHECK(args != nullptr, "cvb_dw_fwd: null args");
  const cvb_dw_fwd_args& a = *args;
  CVB_CHECK(a.B > 0 && a.H > 0 && a.W > 0 && a.C > 0 && a.C % 8 == 0, "cvb_dw_fwd: bad shape B=%d H=%d W=%d C=%d (C %% 8 == 0)", a.B, a.H, a.W, a.C);
  CVB_CHECK(a.stride == 1 || a.stride == 2, "cvb_dw_fwd: stride must be 1 or 2");
  CVB_CHECK(a.X && a.Wt && a.Y && cvb_aligned16(a.X) && cvb_aligned16(a.Y), "cvb_dw_fwd: null / misaligned operand");
  CVB_CHECK(a.x_mode == CVB_A_RAW || ((a.x_mode == CVB_A_AFF || a.x_mode == CVB_A_AFF_SILU) && a.x_p0 && a.x_p1), "cvb_dw_fwd: bad x_mode %d", a.x_mode);
  if (a.col_sum) CVB_CHECK(a.col_sq != nullptr, "cvb_dw_fwd: col_sq missing");
  const int Ho = (a.H - 1) / a.stride + 1, Wo = (a.W - 1) / a.stride + 1;
  int TH, TW; pick_tile(Ho, Wo, a.stride, &TH, &TW);
  const int tiles_h = (Ho + TH - 1) / TH, tiles_w = (Wo + TW - 1) / TW;
  const int IH = (TH - 1) * a.stride + 3, IW = (TW - 1) * a.stride + 3;
  const int buf_bytes = (IH * IW * 128 + 1023) / 1024 * 1024;  // 128B-swizzled TMA destinations are 1024-byte aligned
  size_t smem = (size_t)2 * buf_bytes + 1024;
  const int cblocks = (a.C + CB - 1) / CB;
  int per_img = tiles_h * tiles_w * cblocks;
  int gz = (8 * cvb_num_sms() + per_img - 1) / per_img;  // batch loop inside the CTA: double-buffered TMA + bounded statistics atomics
  if (gz > a.B) gz = a.B;
  if (gz < 1) gz = 1;
  dim3 grid(tiles_h * tiles_w, cblocks, gz);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmX;
  if (cvb_make_tmap_nhwc(&tmX, a.X, a.B, a.H, a.W, a.C, IH, IW, CB, 0)) return 1;
#define CVB_DW_FWD(MODE)                                                                                                  \
  {                                                                                                                      \
    static bool attr = false;                                                                                            \
    if (!attr) { CVB_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; } \
    CVB_CUDA(cvb_launch(dw_fwd_kernel<MODE>, grid, NT, smem, st, tmX, a, Ho, Wo, TH, TW, ilog2(TW), tiles_w, buf_bytes));                   \
  }
  if (a.x_mode == CVB_A_RAW) CVB_DW_FWD(CVB_A_RAW)
  else if (a.x_mode == CVB_A_AFF) CVB_DW_FWD(CVB_A_AFF)
  else CVB_DW_FWD(CVB_A_AFF_SILU)
#undef CVB_DW_FWD
  CVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int cvb_dw_bwd(const cvb_dw_bwd_args* args, cvb_stream_t stream) {
  CVB_CHECK(args != nullptr, "cvb_dw_bwd: null args");
  const cvb_dw_bwd_args& a = *args;
  CVB_CHECK(a.B > 0 && a.H > 0 && a.W > 0 && a.C > 0 && a.C % 8 == 0, "cvb_dw_bwd: bad shape");
  CVB_CHECK(a.stride == 1 || a.stride == 2, "cvb_dw_bwd: stride must be 1 or 2");
  CVB_CHECK(a.DZ && a.X && a.Wt && a.DX && a.dWt && cvb_aligned16(a.DZ) && cvb_aligned16(a.X) && cvb_aligned16(a.DX), "cvb_dw_bwd: null / misaligned operand");
  CVB_CHECK(a.g_mode == CVB_A_RAW || (a.g_mode == CVB_A_BNB && a.Y2 && a.g_p0 && a.g_p1 && a.g_p2), "cvb_dw_bwd: bad g_mode %d", a.g_mode);
  CVB_CHECK(a.x_mode == CVB_A_RAW || ((a.x_mode == CVB_A_AFF || a.x_mode == CVB_A_AFF_SILU) && a.x_p0 && a.x_p1), "cvb_dw_bwd: bad x_mode %d", a.x_mode);
  if (a.col_sum) CVB_CHECK(a.col_sq != nullptr, "cvb_dw_bwd: col_sq missing");
  if (a.stride == 2) CVB_CHECK(a.H % 2 == 0 && a.W % 2 == 0, "cvb_dw_bwd: stride 2 needs even H, W");
  const int Ho = (a.H - 1) / a.stride + 1, Wo = (a.W - 1) / a.stride + 1;
  int TH, TW; pick_tile(Ho, Wo, a.stride, &TH, &TW);
  const int s = a.stride;
  const int tiles_h = (Ho + TH - 1) / TH, tiles_w = (Wo + TW - 1) / TW;
  const int GH = TH + (s == 1 ? 2 : 1), GW = TW + (s == 1 ? 2 : 1);
  const int XH = s * TH + 3 - s, XW = s * TW + 3 - s;
  const int g_bytes = (GH * GW * 128 + 1023) / 1024 * 1024, x_bytes = (XH * XW * 128 + 1023) / 1024 * 1024;
  const int nG = (a.g_mode == CVB_A_BNB) ? 2 : 1;
  size_t smem = (size_t)2 * (nG * g_bytes + x_bytes) + 1024;
  const int cblocks = (a.C + CB - 1) / CB;
  // batch loop inside the CTA (double-buffered TMA, dW / statistics flushed once): one CTA per SM, a few waves
  int per_img = tiles_h * tiles_w * cblocks;
  int want = 4 * cvb_num_sms();
  int gz = (want + per_img - 1) / per_img;
  if (gz > a.B) gz = a.B;
  if (gz < 1) gz = 1;
  dim3 grid(tiles_h * tiles_w, cblocks, gz);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmDZ, tmY2, tmX;
  if (cvb_make_tmap_nhwc(&tmDZ, a.DZ, a.B, Ho, Wo, a.C, GH, GW, CB, 0)) return 1;
  if (cvb_make_tmap_nhwc(&tmY2, a.g_mode == CVB_A_BNB ? a.Y2 : a.DZ, a.B, Ho, Wo, a.C, GH, GW, CB, 0)) return 1;
  if (cvb_make_tmap_nhwc(&tmX, a.X, a.B, a.H, a.W, a.C, XH, XW, CB, 0)) return 1;
#define CVB_DW_BWD(GM, XM)                                                                                                \
  {                                                                                                                      \
    static bool attr = false;                                                                                            \
    if (!attr) { CVB_CUDA(cudaFuncSetAttribute(dw_bwd_kernel<GM, XM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024)); attr = true; } \
    CVB_CUDA(cvb_launch(dw_bwd_kernel<GM, XM>, grid, NTB, smem, st, tmDZ, tmY2, tmX, a, Ho, Wo, TH, TW, ilog2(TW), tiles_w, g_bytes, x_bytes));   \
  }
  if (a.g_mode == CVB_A_RAW) {
    if (a.x_mode == CVB_A_RAW) CVB_DW_BWD(CVB_A_RAW, CVB_A_RAW)
    else if (a.x_mode == CVB_A_AFF) CVB_DW_BWD(CVB_A_RAW, CVB_A_AFF)
    else CVB_DW_BWD(CVB_A_RAW, CVB_A_AFF_SILU)
  } else {
    if (a.x_mode == CVB_A_RAW) CVB_DW_BWD(CVB_A_BNB, CVB_A_RAW)
    else if (a.x_mode == CVB_A_AFF) CVB_DW_BWD(CVB_A_BNB, CVB_A_AFF)
    else CVB_DW_BWD(CVB_A_BNB, CVB_A_AFF_SILU)
  }
#undef CVB_DW_BWD
  CVB_LAUNCH_CHECK();
  return 0;
}
