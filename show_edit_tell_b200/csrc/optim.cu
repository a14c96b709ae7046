// Error reporting, version, and the fused optimizer tail of the train step:
// global-norm clip + Adam over one flat buffer (train() tail, editnet.py:580-581).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "../../include/set_b200.h"
#include "common.cuh"

namespace {
std::mutex g_err_mu;
std::string g_err = "";
}  // namespace

extern "C" void set_record_error(const char* msg) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = msg ? msg : "";
}

extern "C" const char* set_last_error(void) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  static thread_local std::string copy;
  copy = g_err;
  return copy.c_str();
}

extern "C" int set_version(void) { return 100; }

namespace {
std::atomic<long long> g_launches{0};
}
extern "C" void set_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long set_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

namespace set {

void* lib_scratch(int tag, cudaStream_t stream, size_t bytes, bool zero) {
  static std::mutex mu;
  static std::map<std::tuple<int, cudaStream_t, int>, std::pair<void*, size_t>> blocks;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_record_error("lib_scratch: cudaGetDevice failed"); return nullptr; }
  std::lock_guard<std::mutex> lk(mu);
  auto& e = blocks[std::make_tuple(dev, stream, tag)];
  if (e.second >= bytes && e.first) return e.first;
  if (e.first) cudaFree(e.first);   // synchronises the device: no in-flight kernel still reads the old block
  e.first = nullptr; e.second = 0;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); set_record_error("lib_scratch: cudaMalloc failed"); return nullptr; }
  if (zero && cudaMemset(p, 0, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(p); set_record_error("lib_scratch: cudaMemset failed"); return nullptr; }
  e.first = p; e.second = bytes;
  return p;
}

namespace {

// Global-norm clip, stage 1: block b writes the sum of squares of its (fixed) grid-stride share to part[b].  Stage 2
// lives in adam_kernel: every block adds the partials in the same fixed order, so the clip coefficient is a pure
// function of the gradient bits -- data-parallel replicas that hold bit-identical all-reduced gradients stay
// bit-identical (a float atomicAdd across blocks would make the coefficient depend on arrival order).
constexpr int kOptBlocks = 148 * 8;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ part) {
  __shared__ float red[40];
  float s = 0.f;
  const size_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n4; x += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + x);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (size_t x = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n;
       x += (size_t)gridDim.x * blockDim.x)
    s += g[x] * g[x];
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// torch.optim.Adam (defaults, no amsgrad / weight decay) after clip_grad_norm_:
//   g *= min(1, max_norm / (total_norm + 1e-6));  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, size_t n,
                                                   float lr, float b1, float b2, float eps, float bc1, float bc2s,
                                                   float max_norm, float grad_scale,
                                                   const float* __restrict__ count_dev,
                                                   float* __restrict__ scratch) {
  __shared__ float red[40];
  float part = 0.f;
  for (int b = threadIdx.x; b < kOptBlocks; b += blockDim.x) part += scratch[8 + b];   // same order in every block
  const float sumsq = block_sum(part, red);
  // DP: ranks contributed sums; a zero global count (every shard empty) leaves the parameters untouched
  if (count_dev) grad_scale = count_dev[0] > 0.f ? grad_scale / count_dev[0] : 0.f;
  const float total = sqrtf(sumsq) * fabsf(grad_scale);
  const float coef = fminf(max_norm / (total + 1e-6f), 1.0f) * grad_scale;
  if (blockIdx.x == 0 && threadIdx.x == 0) scratch[1] = total;
  const float step = lr / bc1;
  auto upd = [&](float gx, float& mx, float& vx, float& px) {
    const float gv = gx * coef;
    mx = b1 * mx + (1.f - b1) * gv;
    vx = b2 * vx + (1.f - b2) * gv * gv;
    px -= step * mx / (sqrtf(vx) / bc2s + eps);
  };
  // 128-bit accesses when the four buffers allow it (they are views of flat, aligned allocations)
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  const size_t n4 = vec ? (n >> 2) : 0;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n4; x += (size_t)gridDim.x * blockDim.x) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + x);
    float4 m4 = reinterpret_cast<float4*>(m)[x], v4 = reinterpret_cast<float4*>(v)[x], p4 = reinterpret_cast<float4*>(p)[x];
    upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(m)[x] = m4;
    reinterpret_cast<float4*>(v)[x] = v4;
    reinterpret_cast<float4*>(p)[x] = p4;
  }
  for (size_t x = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
    float mx = m[x], vx = v[x], px = p[x];
    upd(g[x], mx, vx, px);
    m[x] = mx; v[x] = vx; p[x] = px;
  }
}

}  // namespace
}  // namespace set

extern "C" int set_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n,
                             int step, float lr, float beta1, float beta2, float eps, float max_norm,
                             float grad_scale, const float* count_dev, float* scratch, void* stream) {
  SET_REQUIRE(params && grads && exp_avg && exp_avg_sq && scratch && step >= 1, "bad args");
  SET_REQUIRE((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "grads must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = set::kOptBlocks;
  set::sumsq_kernel<<<blocks, 256, 0, st>>>(grads, n, scratch + 8);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  set::adam_kernel<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2s,
                                           max_norm, grad_scale, count_dev, scratch);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}
