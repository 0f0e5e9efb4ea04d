// Shared device/host helpers for libcvnets_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cvnets_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcvnets_b200 targets sm_100a (B200) only"
#endif

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---------------------------------------------------------------------------------------------- error handling
void cvb_set_error(const char* fmt, ...);
#define CVB_CHECK(cond, ...)            \
  do {                                  \
    if (!(cond)) {                      \
      cvb_set_error(__VA_ARGS__);       \
      return 1;                         \
    }                                   \
  } while (0)
#define CVB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      (void)cudaGetLastError(); /* reported here: do not leave it for the next caller */ \
      cvb_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)
#define CVB_LAUNCH_CHECK() CVB_CUDA(cudaGetLastError())

static inline bool cvb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int cvb_num_sms();

// ---------------------------------------------------------------------------------------------- small device helpers
// SiLU through ONE special-function op: sigmoid(z) = 0.5 + 0.5 tanh(z/2) with tanh.approx.f32 (MUFU.TANH, max abs error 2^-11), instead of
// ex2 + rcp (two MUFU ops: 16 / clk / SM on B200, which made the SiLU of a 268 M-element tensor cost ~120 us of MUFU pipe alone).  The
// absolute error of silu is <= |z| * 2.5e-4 -- below bf16 resolution of every value that is not itself negligible.  -DCVB_SILU_EXP=1
// restores the exp formulation (A/B and accuracy checks).
#ifndef CVB_SILU_EXP
#define CVB_SILU_EXP 0
#endif
__device__ __forceinline__ float tanh_approx_f(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float z) {
#if CVB_SILU_EXP
  return 1.0f / (1.0f + __expf(-z));
#else
  return fmaf(0.5f, tanh_approx_f(0.5f * z), 0.5f);
#endif
}
__device__ __forceinline__ float silu_f(float z) {
#if CVB_SILU_EXP
  return z / (1.0f + __expf(-z));
#else
  const float h = 0.5f * z;
  return fmaf(h, tanh_approx_f(h), h);
#endif
}
// d/dz [z*sigmoid(z)] = s*(1 + z*(1-s))
__device__ __forceinline__ float silu_grad_f(float z) {
  const float s = sigmoid_f(z);
  return s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf162(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf162(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}
// packed fp32 arithmetic on a register pair (sm_100: FFMA2 / FMUL2 / FADD2): halves the issue slots of pair-wise fp32 math
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{ .reg .b64 a,b,c,d; mov.b64 a,{%2,%3}; mov.b64 b,{%4,%5}; mov.b64 c,{%6,%7}; fma.rn.f32x2 d,a,b,c; mov.b64 {%0,%1}, d; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 a,b,d; mov.b64 a,{%2,%3}; mov.b64 b,{%4,%5}; mul.rn.f32x2 d,a,b; mov.b64 {%0,%1}, d; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 a,b,d; mov.b64 a,{%2,%3}; mov.b64 b,{%4,%5}; add.rn.f32x2 d,a,b; mov.b64 {%0,%1}, d; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// 8 bf16 <-> 8 floats
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf162(u.x), b = unpack_bf162(u.y), c = unpack_bf162(u.z), d = unpack_bf162(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf162(f[0], f[1]); u.y = pack_bf162(f[2], f[3]); u.z = pack_bf162(f[4], f[5]); u.w = pack_bf162(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// streaming (read-once / write-once) 16-byte accesses: do not pollute L1
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// cp.async 16B with zero-fill when !pred (src-size 0)
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gptr), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// apply a per-channel "load mode" to one value
__device__ __forceinline__ float apply_mode(int mode, float x, float p0, float p1) {
  switch (mode) {
    case CVB_A_AFF: return fmaf(p0, x, p1);
    case CVB_A_AFF_SILU: return silu_f(fmaf(p0, x, p1));
    case CVB_A_SILU: return silu_f(x);
    default: return x;
  }
}

// ---------------------------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library is launched with the programmatic-stream-serialization attribute: its CTAs may become resident
// while the previous kernel in the stream is still draining, run their private set-up (mbarrier init, TMEM allocation,
// tensor-map prefetch), and block in pdl_wait() until the previous grid has completed and flushed its memory.  RULES: nothing
// produced by an earlier kernel is read, and no global memory is written, before pdl_wait(); every kernel executes
// pdl_wait() in every thread (so "grid B complete" always implies "grid A complete" along the stream).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int cvb_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t cvb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cvb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------- TMA + mbarrier (sm_90+/sm_100a)
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

// Host: tensor map of a channels-last bf16 feature map [B, H, W, C] with a [1, boxH, boxW, boxC] box (boxC * 2 bytes <= 128),
// optional 128-byte swizzle, zero fill for out-of-bounds elements (the conv halo).
int cvb_make_tmap_nhwc(CUtensorMap* map, const void* base, int B, int H, int W, int C, int boxH, int boxW, int boxC, int swizzle128);
// Host: tensor map of a row-major bf16 matrix [rows, cols] (leading dimension ld elements) with a [box_rows, 32 cols] box and
// 64-byte swizzle: the shared-memory image is exactly the 64-byte-row XOR layout (swz64) the GEMM's ldmatrix addressing uses.
int cvb_make_tmap_2d_k32(CUtensorMap* map, const void* base, int64_t rows, int cols, int ld, int box_rows);
// Same matrix view with a [box_rows, 64 cols] box and 128-byte swizzle: the image is the canonical MN-major SWIZZLE_128B UMMA
// operand layout (8-row x 128-byte atoms) when the ROWS are the reduction dimension (weight-gradient GEMM).
int cvb_make_tmap_2d_c64(CUtensorMap* map, const void* base, int64_t rows, int cols, int ld, int box_rows);

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// make generic-proxy writes/reads of smem visible to the async proxy (TMA) before it overwrites the buffer
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 2-D tiled TMA load: coordinates (col, row)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row)
               : "memory");
}
// 4-D tiled TMA load: coordinates innermost first (c, w, h, b); completes `bytes of the box` on the mbarrier
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h, int b) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(b)
               : "memory");
}
