// cuda_emu.h — TEST INFRASTRUCTURE ONLY.
//
// A small host emulation of the CUDA execution model so that the SIMT kernels and the whole
// host-side lowering / plan executor of libtsb can be exercised by the CPU test-suite in the
// build container (which has no GPU).  Compiled only into tests/emu/_build/libtsb_emu.so
// (-DTSB_HOST_EMU, plain g++); the product library libtsb.so never contains it and the
// package never loads the emulated library: there is no CPU fallback in the product path.
//
// Model: blocks run one after another; the threads of a block are ucontext fibers scheduled
// round-robin; __syncthreads() and the warp collectives yield to the scheduler.
#pragma once
#ifdef TSB_HOST_EMU

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__

struct uint3e { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;  // 0 runnable, 1 at block barrier, 2 finished
};

struct State {
  uint3e threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
  dim3 blockDim, gridDim;
  unsigned char* dyn_smem = nullptr;
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  const std::function<void()>* body = nullptr;
  // warp collectives
  std::vector<uint64_t> wbuf;      // [warp][parity][lane]
  std::vector<int> wcount;         // [warp][parity]
  std::vector<int> wseq;           // [thread] number of collectives entered
};

inline State& S() {
  static State s;
  return s;
}

static const size_t kStack = 256 * 1024;

inline void yield_to_sched() {
  State& s = S();
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void trampoline() {
  State& s = S();
  (*s.body)();
  s.fibers[s.cur].state = 2;
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void set_thread(int t) {
  State& s = S();
  s.cur = t;
  s.threadIdx.x = t % s.blockDim.x;
  s.threadIdx.y = (t / s.blockDim.x) % s.blockDim.y;
  s.threadIdx.z = t / (s.blockDim.x * s.blockDim.y);
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State& s = S();
  s.gridDim = grid;
  s.blockDim = block;
  s.body = &body;
  const int nt = block.x * block.y * block.z;
  if ((int)s.fibers.size() < nt) {
    size_t old = s.fibers.size();
    s.fibers.resize(nt);
    for (size_t i = old; i < (size_t)nt; i++) s.fibers[i].stack = (char*)malloc(kStack);
  }
  std::vector<unsigned char> smem_buf(smem + 64);
  s.dyn_smem = (unsigned char*)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
  const int nw = (nt + 31) / 32;
  s.wbuf.assign((size_t)nw * 2 * 32, 0);
  s.wcount.assign((size_t)nw * 2, 0);
  s.wseq.assign(nt, 0);
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        s.blockIdx = uint3e{bx, by, bz};
        std::fill(s.wcount.begin(), s.wcount.end(), 0);
        std::fill(s.wseq.begin(), s.wseq.end(), 0);
        for (int t = 0; t < nt; t++) {
          Fiber& f = s.fibers[t];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          f.state = 0;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int finished = 0;
        while (finished < nt) {
          int at_barrier = 0;
          finished = 0;
          for (int t = 0; t < nt; t++) {
            Fiber& f = s.fibers[t];
            if (f.state == 0) {
              set_thread(t);
              swapcontext(&s.sched, &f.ctx);
            }
            if (f.state == 1) at_barrier++;
            if (f.state == 2) finished++;
          }
          if (at_barrier > 0 && at_barrier + finished == nt)
            for (int t = 0; t < nt; t++)
              if (s.fibers[t].state == 1) s.fibers[t].state = 0;
        }
      }
  s.body = nullptr;
}

// all-gather of one 64-bit value per lane across the (full) warp of the calling thread
inline void warp_allgather(uint64_t v, uint64_t out[32], int& nlanes) {
  State& s = S();
  const int nt = s.blockDim.x * s.blockDim.y * s.blockDim.z;
  const int t = s.cur, w = t / 32, lane = t % 32;
  nlanes = std::min(32, nt - w * 32);
  const int par = s.wseq[t] & 1;
  s.wseq[t]++;
  s.wbuf[((size_t)w * 2 + par) * 32 + lane] = v;
  s.wcount[w * 2 + par]++;
  while (s.wcount[w * 2 + par] % nlanes != 0 || s.wcount[w * 2 + par] == 0) yield_to_sched();
  for (int l = 0; l < 32; l++) out[l] = l < nlanes ? s.wbuf[((size_t)w * 2 + par) * 32 + l] : 0;
  // no second rendezvous is needed: the buffer of this parity is only reused two collectives
  // on, and a lane can be at most one collective (other parity) ahead of the slowest lane.
}

}  // namespace emu

#define threadIdx (emu::S().threadIdx)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)

static inline void __syncthreads() {
  emu::State& s = emu::S();
  s.fibers[s.cur].state = 1;
  emu::yield_to_sched();
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
  uint64_t v = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    unsigned s4 = (sel >> (4 * i)) & 0xf;
    unsigned byte = (v >> (8 * (s4 & 7))) & 0xff;
    if (s4 & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  uint64_t all[32]; int n; emu::warp_allgather(v, all, n);
  unsigned r = 0; for (int i = 0; i < n; i++) r |= (unsigned)all[i];
  return r;
}
static inline unsigned __reduce_and_sync(unsigned, unsigned v) {
  uint64_t all[32]; int n; emu::warp_allgather(v, all, n);
  unsigned r = 0xffffffffu; for (int i = 0; i < n; i++) r &= (unsigned)all[i];
  return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t all[32]; int n; emu::warp_allgather(pred ? 1 : 0, all, n);
  unsigned r = 0; for (int i = 0; i < n; i++) r |= (unsigned)(all[i] & 1) << i;
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
  uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
  uint64_t all[32]; int n; emu::warp_allgather(bits, all, n);
  int src = (emu::S().cur % 32) ^ lanemask;
  uint64_t b = src < n ? all[src] : bits;
  T r; memcpy(&r, &b, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int srcLane) {
  uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
  uint64_t all[32]; int n; emu::warp_allgather(bits, all, n);
  uint64_t b = all[srcLane & 31];
  T r; memcpy(&r, &b, sizeof(T)); return r;
}
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; *p = o < v ? o : v; return o; }

// ---- runtime API subset -------------------------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)(uintptr_t)1; return cudaSuccess; }
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  static uintptr_t next = 16;  // distinct fake handles (nothing dereferences them)
  *s = (void*)(next += 16);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

#define TSB_DYN_SMEM(name) unsigned char* name = emu::S().dyn_smem
#define TSB_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define TSB_LAUNCH_PDL TSB_LAUNCH
#define TSB_PDL_TRIGGER() ((void)0)
#define TSB_PDL_WAIT() ((void)0)

#else  // real CUDA

#define TSB_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define TSB_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)

// Programmatic dependent launch (sm_90+): a kernel launched with TSB_LAUNCH_PDL may start while the kernel
// before it in the stream is still draining.  Such a kernel builds its shared-memory tables from its
// parameters, then TSB_PDL_WAIT()s (griddepcontrol.wait: the preceding grid has completed and its writes are
// visible) before it reads or writes global memory; TSB_PDL_TRIGGER() at its top lets its own successor do the
// same.  Kernels launched with plain TSB_LAUNCH keep full stream order on both sides.  TSB_PDL=0 disables it.
#define TSB_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define TSB_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

namespace tsb {
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("TSB_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
}  // namespace tsb
#define TSB_LAUNCH_PDL(kern, grid, block, smem, stream, ...) \
  ::tsb::launch_pdl(kern, (grid), (block), (smem), (cudaStream_t)(stream), __VA_ARGS__)

#endif
