// Shared device helpers for the EditNet/DCNet decode path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#define SET_OK 0
#define SET_ERR_ARG 1
#define SET_ERR_CUDA 2
#define SET_ERR_WORKSPACE 3

extern "C" void set_record_error(const char* msg);
extern "C" void set_count_launch(int n);  // bumps the library-wide kernel-launch counter

#define SET_CHECK_CUDA(expr)                                                       \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      char _buf[512];                                                              \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,   \
               cudaGetErrorString(_e));                                            \
      set_record_error(_buf);                                                      \
      return SET_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define SET_REQUIRE(cond, msg)                                                     \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      char _buf[512];                                                              \
      snprintf(_buf, sizeof(_buf), "%s:%d: requirement failed: %s (%s)", __FILE__, \
               __LINE__, #cond, msg);                                              \
      set_record_error(_buf);                                                      \
      return SET_ERR_ARG;                                                          \
    }                                                                              \
  } while (0)

#define SET_PROPAGATE(expr)       \
  do {                            \
    int _r = (expr);              \
    if (_r != SET_OK) return _r;  \
  } while (0)

namespace set {

// Library-owned device scratch (split-K slabs + arrival counters of the tensor-core GEMM, the d-alpha buffer between
// the two attention-backward kernels, ...).  One allocation per (device, stream, tag): two streams or two devices of one
// process never share a buffer, so the C ABI may be driven from several streams / threads / devices at once.  Grows on
// demand (cudaFree of the old block synchronises the device); `zero` clears a fresh block.  nullptr on failure
// (set_last_error() says why).
enum ScratchTag : int { kScratchTcSlabs = 1, kScratchTcCounters = 2, kScratchAttnDal = 3, kScratchStepBarrier = 4,
                        kScratchStepSlabs = 5, kScratchStepMaps = 6 };
void* lib_scratch(int tag, cudaStream_t stream, size_t bytes, bool zero);

// ---------------------------------------------------------------------------------
// Programmatic dependent launch.  The decode step is a chain of short dependent kernels; each
// kernel of the chain is launched with the programmatic-serialization attribute, signals its
// dependents at entry (pdl_trigger) and blocks (pdl_wait) only where it first touches data the
// previous kernels produce.  What precedes the wait -- barrier/TMEM set-up and, in the GEMM, the
// TMA stream of the constant weight operand -- overlaps the tail of the previous kernels.
// Rules: (1) every kernel launched through launch_chain() executes pdl_wait() before it reads or
// writes anything a predecessor touches; (2) kernels that WRITE weights (optimizer, transposes)
// never call pdl_trigger(), so a prefetching GEMM can never run beside them.
// ---------------------------------------------------------------------------------
extern int g_pdl;   // 1: chain launches carry the attribute (default; SET_PDL=0 disables)

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// launch_chain with a thread-block cluster of `cluster` CTAs along x (grid.x must be a multiple)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

constexpr float kNegFill = -1e10f;  // editnet.py:374 masked_fill value

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

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

// Block-wide sum; `scratch` holds >= 33 floats.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) scratch[32] = r;
  }
  __syncthreads();
  return scratch[32];
}

// ---------------------------------------------------------------------------------
// Counter-based dropout.  One Philox4x32-10 call yields 128 keep bits; element `idx`
// of dropout site `site` at `seed` uses counter (idx >> 7) and bit (idx & 127).  The
// forward and backward kernels regenerate the same bits, so no mask is ever stored;
// set_dropout_keep_mask() materialises them for parity tests.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

enum DropSite : uint32_t { kSiteEnc = 1, kSiteEmb = 2, kSiteVis = 3, kSiteFc = 4, kSiteSample = 5 };

// 128 keep bits for elements [128*blk, 128*blk+127]
__device__ __forceinline__ uint4 drop_bits128(uint64_t seed, uint32_t site, uint64_t blk) {
  uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), site, 0x5e7b200u);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return philox4x32_10(ctr, key);
}
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint4 b = drop_bits128(seed, site, idx >> 7);
  const uint32_t bit = (uint32_t)idx & 127u;
  const uint32_t w = bit < 64 ? (bit < 32 ? b.x : b.y) : (bit < 96 ? b.z : b.w);
  return (w >> (bit & 31u)) & 1u;
}
// four keep bits for elements idx..idx+3 (idx % 4 == 0)
__device__ __forceinline__ uint32_t drop_keep4(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint4 b = drop_bits128(seed, site, idx >> 7);
  const uint32_t bit = (uint32_t)idx & 127u;
  const uint32_t w = bit < 64 ? (bit < 32 ? b.x : b.y) : (bit < 96 ? b.z : b.w);
  return (w >> (bit & 31u)) & 0xFu;
}

// uniform in [0,1) for sampling; one per (site, idx)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint4 b = drop_bits128(seed, site, idx);
  return (float)(b.x >> 8) * (1.0f / 16777216.0f);
}

}  // namespace set
