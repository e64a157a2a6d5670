// k_generic.cuh — generic strided contraction / rewire kernel and small elementwise kernels.
//
// The generic kernel is the on-device parity reference: any
// extents, any strides, repeated labels, dropped labels.  It evaluates
//     out[sum_f i_f*so_f] (op)= sum_{reduce dims} lhs[sum i*sl] * rhs[sum i*sr]
// with one thread per output iteration point.  It backs FinRel._contract2 for non-binary
// extents and tiny tensors (tensorsat/lang/fin_rel.py:255-288), rewire_array
// (tensorsat/_utils/misc.py:104-123), _getitem (:342-343), _stack (:304-308) and leaf
// preparation inside plans.
#pragma once
#include "tsb_common.h"

namespace tsb {

constexpr int kGenMaxDims = 64;

struct GenDesc {
  int32_t n_free, n_red;
  int32_t ext[kGenMaxDims];  // free dims first (last free dim iterates fastest), then reduce dims
  int64_t sl[kGenMaxDims], sr[kGenMaxDims], so[kGenMaxDims];
  int64_t n_out_iter, n_red_iter;
  int32_t dl, dr, dout;  // tsb_dtype
  int32_t clamp;         // 1: store (acc != 0)
  int32_t has_rhs;
  int32_t accumulate;    // 0 store, 1 combine with existing value (or for clamp, + otherwise)
};

TSB_HD int64_t load_as_i64(const void* p, int dt, int64_t idx) {
  if (dt == TSB_U8) return ((const uint8_t*)p)[idx];
  if (dt == TSB_I32) return ((const int32_t*)p)[idx];
  return ((const int64_t*)p)[idx];
}
TSB_HD void store_from_i64(void* p, int dt, int64_t idx, int64_t v) {
  if (dt == TSB_U8) ((uint8_t*)p)[idx] = (uint8_t)v;
  else if (dt == TSB_I32) ((int32_t*)p)[idx] = (int32_t)v;
  else ((int64_t*)p)[idx] = v;
}

static __global__ void __launch_bounds__(256)
k_generic(const __grid_constant__ GenDesc d, const void* __restrict__ lhs,
          const void* __restrict__ rhs, void* __restrict__ out) {
  TSB_PDL_TRIGGER();
  TSB_PDL_WAIT();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < d.n_out_iter; it += stride) {
    int64_t ol = 0, orr = 0, oo = 0, rem = it;
    for (int f = d.n_free - 1; f >= 0; f--) {
      const int64_t e = d.ext[f];
      const int64_t i = rem % e;
      rem /= e;
      ol += i * d.sl[f];
      orr += i * d.sr[f];
      oo += i * d.so[f];
    }
    int64_t acc = 0;
    for (int64_t r = 0; r < d.n_red_iter; r++) {
      int64_t pl = ol, pr = orr, rr = r;
      for (int q = d.n_free + d.n_red - 1; q >= d.n_free; q--) {
        const int64_t e = d.ext[q];
        const int64_t i = rr % e;
        rr /= e;
        pl += i * d.sl[q];
        pr += i * d.sr[q];
      }
      const int64_t a = load_as_i64(lhs, d.dl, pl);
      const int64_t b = d.has_rhs ? load_as_i64(rhs, d.dr, pr) : 1;
      acc += a * b;
    }
    if (d.accumulate) acc += load_as_i64(out, d.dout, oo);
    if (d.clamp) acc = acc != 0;
    store_from_i64(out, d.dout, oo, acc);
  }
}

// ---- batched small ops ---------------------------------------------------------------------
// One CTA per op: the thousands of leaf preparations and rank <= ~14 steps at the bottom (and the
// tail) of a contraction tree are launch-latency bound, so every wave of mutually
// independent small ops of a plan is ONE launch.  Same arithmetic as k_generic; operands are
// addressed through a pointer table (arena base, result buffer, leaf tensors) so that the op list
// is built once per plan.
constexpr int kSmallMaxDims = 12;
constexpr int kSmallThreads = 128;

struct SmallOp {
  int32_t n_free, n_red, n_out_iter, n_red_iter;
  int32_t ext[kSmallMaxDims], sl[kSmallMaxDims], sr[kSmallMaxDims], so[kSmallMaxDims];
  int32_t lhs_buf, rhs_buf, out_buf;  // indices into the pointer table
  int32_t zero_first;                 // number of output elements to clear first (repeated labels)
  int64_t lhs_off, rhs_off, out_off;  // byte offsets
  int8_t dl, dr, dout, clamp, has_rhs, binary /* 1: every extent is 2 */, pad1, pad2;
};

// one small op by the threads of one CTA
// (part, n_parts): the outputs of one op can be shared out over several CTAs (the larger ops of a wave); an op that
// clears its result first (repeated labels) is done by part 0 alone, its barrier being a block barrier
TSB_D void small_op_run(const SmallOp& d, void* const* __restrict__ ptrs, int part = 0, int n_parts = 1) {
  const void* lhs = (const uint8_t*)ptrs[d.lhs_buf] + d.lhs_off;
  const void* rhs = d.has_rhs ? (const void*)((const uint8_t*)ptrs[d.rhs_buf] + d.rhs_off) : nullptr;
  void* out = (uint8_t*)ptrs[d.out_buf] + d.out_off;
  if (d.zero_first) {
    if (part != 0) return;
    n_parts = 1;
    for (int i = threadIdx.x; i < d.zero_first; i += blockDim.x) store_from_i64(out, d.dout, i, 0);
    __syncthreads();
  }
  const int first = threadIdx.x + part * blockDim.x, step = blockDim.x * n_parts;
  if (d.binary) {
    // every extent is 2 (all BASELINE networks): index digits are bits, no integer division
    for (int it = first; it < d.n_out_iter; it += step) {
      int ol = 0, orr = 0, oo = 0;
      for (int f = d.n_free - 1, q = 0; f >= 0; f--, q++) {
        const int i = (it >> q) & 1;
        ol += i * d.sl[f];
        orr += i * d.sr[f];
        oo += i * d.so[f];
      }
      int64_t acc = 0;
      for (int r = 0; r < d.n_red_iter; r++) {
        int pl = ol, pr = orr;
        for (int q = d.n_free + d.n_red - 1, z = 0; q >= d.n_free; q--, z++) {
          const int i = (r >> z) & 1;
          pl += i * d.sl[q];
          pr += i * d.sr[q];
        }
        const int64_t a = load_as_i64(lhs, d.dl, pl);
        const int64_t b = d.has_rhs ? load_as_i64(rhs, d.dr, pr) : 1;
        acc += a * b;
      }
      if (d.clamp) acc = acc != 0;
      store_from_i64(out, d.dout, oo, acc);
    }
    return;
  }
  for (int it = first; it < d.n_out_iter; it += step) {
    int ol = 0, orr = 0, oo = 0, rem = it;
    for (int f = d.n_free - 1; f >= 0; f--) {
      const int e = d.ext[f];
      const int i = rem % e;
      rem /= e;
      ol += i * d.sl[f];
      orr += i * d.sr[f];
      oo += i * d.so[f];
    }
    int64_t acc = 0;
    for (int r = 0; r < d.n_red_iter; r++) {
      int pl = ol, pr = orr, rr = r;
      for (int q = d.n_free + d.n_red - 1; q >= d.n_free; q--) {
        const int e = d.ext[q];
        const int i = rr % e;
        rr /= e;
        pl += i * d.sl[q];
        pr += i * d.sr[q];
      }
      const int64_t a = load_as_i64(lhs, d.dl, pl);
      const int64_t b = d.has_rhs ? load_as_i64(rhs, d.dr, pr) : 1;
      acc += a * b;
    }
    if (d.clamp) acc = acc != 0;
    store_from_i64(out, d.dout, oo, acc);
  }
}

// The descriptor is staged in shared memory first: read through a global pointer every field is re-fetched inside the
// index loops (the compiler must assume the result stores alias it), which made these microsecond-sized ops several
// times slower than their arithmetic.
TSB_D void small_op_stage(const SmallOp* __restrict__ src, SmallOp& dst) {
  static_assert(sizeof(SmallOp) % 4 == 0, "SmallOp is copied word by word");
  for (int i = threadIdx.x; i < (int)(sizeof(SmallOp) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  __syncthreads();
}

static __global__ void __launch_bounds__(kSmallThreads)
k_small_batch(const SmallOp* __restrict__ ops, void* const* __restrict__ ptrs) {
  __shared__ SmallOp sd;
  TSB_PDL_TRIGGER();
  small_op_stage(ops + blockIdx.x, sd);   // descriptors are plan constants: staged before the dependency wait
  TSB_PDL_WAIT();
  small_op_run(sd, ptrs, (int)blockIdx.y, (int)gridDim.y);   // gridDim.y CTAs share the outputs of every op of the wave
}

// A run of consecutive waves that hold one or two small ops each (the vine at the tail of an elimination-order
// tree, where every step depends on the one before): ONE launch of one wide CTA that executes the ops in order,
// a block barrier between them (global writes of a CTA are visible to it after __syncthreads), instead of one
// launch per wave.
constexpr int kChainThreads = 1024;

static __global__ void __launch_bounds__(kChainThreads)
k_small_chain(const SmallOp* __restrict__ ops, int n_ops, void* const* __restrict__ ptrs) {
  __shared__ SmallOp sd[2];
  TSB_PDL_TRIGGER();
  small_op_stage(ops, sd[0]);
  TSB_PDL_WAIT();
  for (int i = 0; i < n_ops; i++) {
    if (i + 1 < n_ops) {   // next descriptor into the other buffer while this op runs (visible after the barrier below)
      for (int w = threadIdx.x; w < (int)(sizeof(SmallOp) / 4); w += blockDim.x)
        reinterpret_cast<uint32_t*>(&sd[(i + 1) & 1])[w] = reinterpret_cast<const uint32_t*>(ops + i + 1)[w];
    }
    small_op_run(sd[i & 1], ptrs);
    __syncthreads();
  }
}

// ---- elementwise -------------------------------------------------------------------------

// out = a | b (u8 0/1, vectorised) — FinRel._union (tensorsat/lang/fin_rel.py:345-346)
static __global__ void __launch_bounds__(256)
k_or_u8(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, int64_t n) {
  const int64_t nv = n >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t v = t0; v < nv; v += stride) {
    uint4 x = ((const uint4*)a)[v], y = ((const uint4*)b)[v];
    ((uint4*)out)[v] = make_uint4(x.x | y.x, x.y | y.y, x.z | y.z, x.w | y.w);
  }
  for (int64_t i = (nv << 4) + t0; i < n; i += stride) out[i] = a[i] | b[i];
}

template <typename T>
__global__ void __launch_bounds__(256)
k_add(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] + b[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_fill(T* __restrict__ out, int64_t n, T v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

// ones on the diagonal of a (dim ** num_ports) tensor that was zero-filled — FinRel._spider (:290-298)
template <typename T>
__global__ void k_diag_ones(T* __restrict__ out, int dim, int64_t diag_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) out[(int64_t)i * diag_stride] = (T)1;
}

static __global__ void __launch_bounds__(256)
k_convert(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt, int64_t n, int clamp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t v = load_as_i64(src, sdt, i);
    if (clamp) v = v != 0;
    store_from_i64(dst, ddt, i, v);
  }
}

// flag (pre-set to 1) is cleared if any element differs from 1 — FinRel._as_scalar (:472-473)
static __global__ void __launch_bounds__(256)
k_all_ones(const void* __restrict__ p, int dt, int64_t n, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool ok = true;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    ok = ok && (load_as_i64(p, dt, i) == 1);
  if (!ok) *flag = 0;
}

// flag (pre-set to 1) is cleared if any element is neither 0 nor 1: the validation of FinRel.__new__
// (tensorsat/lang/fin_rel.py:119-121, "Values in a Boolean tensor must be 0 or 1.") for device-resident leaves
static __global__ void __launch_bounds__(256)
k_all_binary(const void* __restrict__ p, int dt, int64_t n, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool ok = true;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    ok = ok && ((uint64_t)load_as_i64(p, dt, i) <= 1);
  if (!ok) *flag = 0;
}

}  // namespace tsb
