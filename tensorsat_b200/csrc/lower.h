// lower.h — host-side lowering of box operations to kernel launches.
//
// Everything here is index bookkeeping: which label sits on which address bit of which
// operand, which kernel can serve a step, and how its descriptor is filled.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <set>

#include "k_generic.cuh"
#include "k_merge.cuh"
#include "k_permute.cuh"

namespace tsb {

// ---------------------------------------------------------------------------------------------
// Context: device, caching arena, launch accounting
// ---------------------------------------------------------------------------------------------
// Caching device allocator (north-star item (d)).  Blocks are cached by size; a released block remembers the
// stream its last user ran on and an event recorded at release, and a later allocation on a DIFFERENT stream
// first makes that stream wait for the event, so the host-side release of a block whose kernels are still in
// flight is safe whatever stream asks next.  A mutex guards the bookkeeping.
struct Arena {
  struct FreeBlock {
    void* ptr;
    cudaStream_t stream;  // stream the block's last user was enqueued on
    cudaEvent_t event;    // recorded on that stream at release (nullptr: nothing in flight / unknown)
  };
  std::multimap<size_t, FreeBlock> free_blocks;  // size -> block
  std::map<void*, size_t> live;                  // handed out
  std::map<void*, size_t> owned;                 // every block obtained from the driver
  std::vector<cudaEvent_t> event_pool;
  std::mutex mu;
  tsb_arena_stats_t st{};

  static size_t round_up(size_t n) {
    const size_t q = n < (size_t(1) << 20) ? 512 : (size_t(1) << 20);
    return n == 0 ? q : (n + q - 1) / q * q;
  }
  // `stream`: where the caller will use the block; kNoStream = the caller orders its own work (C-ABI callers)
  static cudaStream_t no_stream() { return (cudaStream_t)(intptr_t)-1; }
  void* alloc(size_t bytes, cudaStream_t stream = no_stream()) {
    std::lock_guard<std::mutex> lock(mu);
    const size_t need = round_up(bytes);
    auto it = free_blocks.lower_bound(need);
    // reuse a cached block unless it would waste more than half of it
    if (it != free_blocks.end() && it->first <= need * 2 + (size_t(1) << 20)) {
      FreeBlock b = it->second;
      const size_t size = it->first;
      free_blocks.erase(it);
      if (b.event) {
        if (b.stream != stream) {
          // the previous user's work may still be running on another stream: order this one after it
          // (a caller that names no stream gets a host-side wait instead)
          if (stream == no_stream()) cudaEventSynchronize(b.event);
          else cudaStreamWaitEvent(stream, b.event, 0);
          st.cross_stream_waits++;
        }
        event_pool.push_back(b.event);
      }
      live[b.ptr] = size;
      st.in_use_bytes += size;
      st.cache_hits++;
      st.peak_in_use_bytes = std::max(st.peak_in_use_bytes, st.in_use_bytes);
      return b.ptr;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {
      trim_locked();
      e = cudaMalloc(&p, need);
    }
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("arena: cudaMalloc of %zu bytes failed: %s", need, cudaGetErrorString(e));
      throw Error{TSB_ERR_OOM};
    }
    owned[p] = need;
    live[p] = need;
    st.reserved_bytes += need;
    st.in_use_bytes += need;
    st.device_allocs++;
    st.peak_in_use_bytes = std::max(st.peak_in_use_bytes, st.in_use_bytes);
    return p;
  }
  // `stream`: where the block's last user was enqueued (an event is recorded there); no_stream(): the caller
  // guarantees that nothing using the block is still in flight
  void release(void* p, cudaStream_t stream = no_stream()) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = live.find(p);
    TSB_CHECK(it != live.end(), TSB_ERR_INVALID, "arena: release of a pointer the arena does not own");
    FreeBlock b{p, stream, nullptr};
    if (stream != no_stream()) {
      if (event_pool.empty()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) event_pool.push_back(e);
      }
      if (!event_pool.empty()) {
        b.event = event_pool.back();
        event_pool.pop_back();
        cudaEventRecord(b.event, stream);
      }
    }
    st.in_use_bytes -= it->second;
    free_blocks.emplace(it->second, b);
    live.erase(it);
  }
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    trim_locked();
  }
  void trim_locked() {
    for (auto& kv : free_blocks) {
      if (kv.second.event) {
        cudaEventSynchronize(kv.second.event);
        event_pool.push_back(kv.second.event);
      }
      cudaFree(kv.second.ptr);
      st.reserved_bytes -= kv.first;
      owned.erase(kv.second.ptr);
    }
    free_blocks.clear();
  }
  ~Arena() {
    for (auto& kv : free_blocks) if (kv.second.event) cudaEventDestroy(kv.second.event);
    for (auto e : event_pool) cudaEventDestroy(e);
    for (auto& kv : owned) cudaFree(kv.first);
  }
};

// temporaries of a captured execution: owned by the graph (replays use the same addresses), recycled within the capture
struct HeldBlocks {
  std::vector<void*> owned;
  std::multimap<size_t, void*> free;
};
// Capture state of the CALLING THREAD (a capture runs synchronously inside one tsb_plan_execute call): where the
// temporaries of the launches go, and where their per-kernel counts are recorded.  Thread-local, so that another
// thread's direct execution on the same context is not diverted into somebody's graph.
struct CaptureState {
  HeldBlocks* hold = nullptr;
  std::map<std::string, int64_t>* counts = nullptr;
};
CaptureState& capture_state();   // tsb_ops.cu

}  // namespace tsb

struct tsb_ctx {
  int refs = 1;            // the creator's reference + one per live plan; freed when it drops to zero
  bool gemm_attr_set = false;
  int device = 0;
  int sm_count = 148;
  int64_t launches = 0;
  // per-kernel launch counts since tsb_ctx_trace(ctx, 1) (tests and smoke() prove which kernels ran)
  bool trace = false;
  std::map<std::string, int64_t> kcount;
  void note(const char* kernel) {
    launches++;
    if (trace) kcount[kernel]++;
    if (auto* c = tsb::capture_state().counts) (*c)[kernel]++;
  }
  tsb::Arena arena;
  int* d_flag = nullptr;  // scratch for all_ones
  uint8_t* d_one = nullptr;  // the u8 scalar 1
};

namespace tsb {

void ctx_unref(tsb_ctx* ctx);  // tsb_ops.cu

// RAII temp buffer from the arena for work enqueued on `stream`.  It is released on the host at scope exit while
// its kernels may still be in flight: the arena orders any later user on another stream after them (event recorded
// at release).  While a plan execution is being captured into a CUDA graph (capture_state().hold != nullptr) the buffer is
// handed to the graph's owner instead: a replay uses the same addresses, so the memory must stay the graph's.
struct Temp {
  tsb_ctx* ctx;
  cudaStream_t stream;
  void* p = nullptr;
  size_t size = 0;
  HeldBlocks* held = nullptr;   // set when the buffer belongs to a capture
  Temp(tsb_ctx* c, size_t bytes, cudaStream_t s) : ctx(c), stream(s) {
    if (HeldBlocks* hold = capture_state().hold) {
      // capture: the launches are recorded in stream order, so a temporary released earlier in the capture can serve a
      // later one; the graph ends up owning each distinct block once
      held = hold;
      size = tsb::Arena::round_up(bytes);
      auto it = hold->free.lower_bound(size);
      if (it != hold->free.end() && it->first <= size * 2 + (size_t(1) << 20)) {
        p = it->second;
        size = it->first;
        hold->free.erase(it);
      } else {
        p = c->arena.alloc(bytes, tsb::Arena::no_stream());
        hold->owned.push_back(p);
      }
      return;
    }
    p = c->arena.alloc(bytes, s);
  }
  ~Temp() {
    if (!p) return;
    if (held) held->free.emplace(size, p);
    else ctx->arena.release(p, stream);
  }
  Temp(const Temp&) = delete;
  Temp& operator=(const Temp&) = delete;
};

inline int grid_for(tsb_ctx* ctx, int64_t work_items, int threads, int per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// Labelled dims: the common currency of the lowering.
// ---------------------------------------------------------------------------------------------
struct LDim {
  int32_t label;
  int64_t ext;
  int64_t sl = 0, sr = 0, so = 0;  // element strides (0 = absent)
  bool in_l = false, in_r = false, in_o = false;
};

// Builds the generic-kernel descriptor from labelled dims (free = in_o).  Free dims iterate in
// decreasing output stride so that consecutive threads write consecutive addresses.
inline GenDesc make_gen_desc(std::vector<LDim> dims, int dl, int dr, int dout, bool clamp,
                             bool has_rhs, bool accumulate) {
  std::vector<LDim> fr, rd;
  for (auto& d : dims) {
    if (d.ext == 1) continue;
    (d.in_o ? fr : rd).push_back(d);
  }
  std::stable_sort(fr.begin(), fr.end(), [](const LDim& a, const LDim& b) { return a.so > b.so; });
  std::stable_sort(rd.begin(), rd.end(), [](const LDim& a, const LDim& b) {
    return a.sl != b.sl ? a.sl > b.sl : a.sr > b.sr;
  });
  auto merge = [](std::vector<LDim>& v) {
    std::vector<LDim> o;
    for (auto& d : v) {
      if (!o.empty()) {
        LDim& p = o.back();  // p is the outer dim, d the inner one
        if (p.sl == d.sl * d.ext && p.sr == d.sr * d.ext && p.so == d.so * d.ext) {
          p.ext *= d.ext;
          p.sl = d.sl; p.sr = d.sr; p.so = d.so;
          continue;
        }
      }
      o.push_back(d);
    }
    v.swap(o);
  };
  merge(fr);
  merge(rd);
  TSB_CHECK((int)(fr.size() + rd.size()) <= kGenMaxDims, TSB_ERR_UNSUPPORTED,
            "generic kernel: too many non-mergeable dimensions (%zu)", fr.size() + rd.size());
  GenDesc g{};
  g.n_free = (int)fr.size();
  g.n_red = (int)rd.size();
  g.n_out_iter = 1;
  g.n_red_iter = 1;
  int i = 0;
  for (auto& d : fr) {
    TSB_CHECK(d.ext <= INT32_MAX, TSB_ERR_UNSUPPORTED, "dimension too large");
    g.ext[i] = (int32_t)d.ext; g.sl[i] = d.sl; g.sr[i] = d.sr; g.so[i] = d.so;
    g.n_out_iter *= d.ext; i++;
  }
  for (auto& d : rd) {
    TSB_CHECK(d.ext <= INT32_MAX, TSB_ERR_UNSUPPORTED, "dimension too large");
    g.ext[i] = (int32_t)d.ext; g.sl[i] = d.sl; g.sr[i] = d.sr; g.so[i] = 0;
    g.n_red_iter *= d.ext; i++;
  }
  g.dl = dl; g.dr = dr; g.dout = dout;
  g.clamp = clamp; g.has_rhs = has_rhs; g.accumulate = accumulate;
  return g;
}

inline void launch_generic(tsb_ctx* ctx, const GenDesc& g, const void* lhs, const void* rhs, void* out,
                           cudaStream_t stream) {
  // the generic kernel is the slow, fully general path: refuse work that would run for minutes
  TSB_CHECK((double)g.n_out_iter * (double)g.n_red_iter <= 68719476736.0 /* 2^36 */, TSB_ERR_UNSUPPORTED,
            "step with %lld outputs x %lld summed terms has no fast kernel (non-binary extents or repeated wires?)",
            (long long)g.n_out_iter, (long long)g.n_red_iter);
  const int grid = grid_for(ctx, g.n_out_iter, 256);
  TSB_LAUNCH_PDL(k_generic, dim3(grid), dim3(256), 0, stream, g, lhs, rhs, out);
  ctx->note("k_generic");
  TSB_CUDA(cudaGetLastError());
}

// Appends runs mapping input bit `in` -> output bit `out` (merging with the previous run).
inline void runs_add(BitRuns& r, int in, int out) {
  if (r.n > 0) {
    const int k = r.n - 1;
    if (r.in_shift[k] + r.len[k] == in && r.out_shift[k] + r.len[k] == out) {
      r.len[k]++;
      return;
    }
  }
  TSB_CHECK(r.n < kMaxRuns, TSB_ERR_UNSUPPORTED, "index map needs more than %d bit runs", kMaxRuns);
  r.in_shift[r.n] = (uint8_t)in;
  r.len[r.n] = 1;
  r.out_shift[r.n] = (uint8_t)out;
  r.n++;
}

// ---------------------------------------------------------------------------------------------
// Bit-level tensors: a tensor with power-of-two extents is a list of bit labels, MSB first.
// ---------------------------------------------------------------------------------------------
using BitLabels = std::vector<int64_t>;  // label * 64 + bit-within-label, most significant first

inline bool all_pow2(const int32_t* ext, int n) {
  for (int i = 0; i < n; i++)
    if (!is_pow2(ext[i])) return false;
  return true;
}

// wires + extents (pow2) -> bit labels, MSB first
inline BitLabels bit_labels(const int32_t* wires, int rank, const int32_t* wire_ext) {
  BitLabels b;
  for (int i = 0; i < rank; i++) {
    const int nb = ilog2(wire_ext[wires[i]]);
    for (int q = nb - 1; q >= 0; q--) b.push_back((int64_t)wires[i] * 64 + q);
  }
  return b;
}

inline std::map<int64_t, int> positions(const BitLabels& b) {
  std::map<int64_t, int> p;
  const int n = (int)b.size();
  for (int i = 0; i < n; i++) p[b[i]] = n - 1 - i;  // address bit
  return p;
}

inline bool has_dups(const BitLabels& b) {
  std::set<int64_t> s(b.begin(), b.end());
  return s.size() != b.size();
}

// ---- permutation lowering ---------------------------------------------------------------
// src_bits / dst_bits: the same label set in two orders (MSB first); esz_log2 extra identical low bits.
// Returns false if the fast kernel cannot serve it (caller falls back to the generic kernel).
inline bool make_perm_desc(const BitLabels& src_bits, const BitLabels& dst_bits, int esz_log2, PermDesc& d, int& G,
                           PermT4Desc* t4 = nullptr, bool* use_t4 = nullptr) {
  const int r = (int)src_bits.size() + esz_log2;  // in byte-address bits
  if (r < 9 || r > 62) return false;
  // label at each byte-address bit of src/dst (element-size bits get pseudo labels -1-q)
  std::vector<int64_t> s(r), t(r);
  for (int q = 0; q < esz_log2; q++) s[q] = t[q] = -1 - q;
  for (int i = 0; i < (int)src_bits.size(); i++) s[r - 1 - i] = src_bits[i];
  for (int i = 0; i < (int)dst_bits.size(); i++) t[r - 1 - i] = dst_bits[i];
  std::map<int64_t, int> spos, dpos;
  for (int i = 0; i < r; i++) { spos[s[i]] = i; dpos[t[i]] = i; }
  int g = 0;
  while (g < 4 && g < r && s[g] == t[g]) g++;
  G = 1 << g;
  const int tb = std::min(r, 14);
  const int a = std::min(7, r), b = std::min(7, r);
  std::set<int64_t> T;
  for (int i = 0; i < a; i++) T.insert(s[i]);
  for (int i = 0; i < b; i++) T.insert(t[i]);
  for (int i = 0; (int)T.size() < tb && i < r; i++) {
    T.insert(t[i]);
    if ((int)T.size() < tb) T.insert(s[i]);
  }
  const int tbits = (int)T.size();
  if (tbits > kPermMaxTileBits) return false;
  // tile-local bit index of each tile label in source order / destination order
  std::vector<int64_t> ts, td;
  for (int i = 0; i < r; i++) if (T.count(s[i])) ts.push_back(s[i]);
  for (int i = 0; i < r; i++) if (T.count(t[i])) td.push_back(t[i]);
  std::map<int64_t, int> tspos, tdpos;
  for (int i = 0; i < tbits; i++) { tspos[ts[i]] = i; tdpos[td[i]] = i; }
  d = PermDesc{};
  d.g = g; d.tbits = tbits; d.a_bits = a; d.b_bits = b;
  for (int i = a; i < tbits; i++) runs_add(d.src_hi, i - a, spos[ts[i]]);
  for (int i = b; i < tbits; i++) runs_add(d.dst_hi, i - b, dpos[td[i]]);
  for (int i = 0; i < tbits; i++) runs_add(d.d2s, i, tspos[td[i]]);
  int ci = 0;
  BitRuns cs{}, cd{};
  // CTA index bits follow destination order so that neighbouring CTAs write neighbouring lines;
  // the source-side map is generally not monotone, so its runs are added one bit at a time.
  for (int i = 0; i < r; i++) {
    if (T.count(t[i])) continue;
    runs_add(cd, ci, i);
    runs_add(cs, ci, spos[t[i]]);
    ci++;
  }
  if (cs.n > 32 || cd.n > 32) return false;  // one run per lane in the kernel
  d.cta_src = cs; d.cta_dst = cd;
  d.n_tiles = int64_t(1) << (r - tbits);
  if (t4 && use_t4) {
    *use_t4 = false;
    // 4x4 register transposes: the two lowest source bits are not among the four lowest destination bits
    if (g == 0 && tbits >= 8) {
      const int a0 = tdpos[ts[0]], a1 = tdpos[ts[1]];
      if (a0 >= 4 && a1 >= 4) {
        *t4 = PermT4Desc{};
        t4->p = d;
        t4->a_dst[0] = a0; t4->a_dst[1] = a1;
        int ui = 0;
        for (int i = 4; i < tbits; i++) {
          if (i == a0 || i == a1) continue;
          runs_add(t4->u2jd, ui, i);
          ui++;
        }
        *use_t4 = true;
      }
    }
  }
  return true;
}

inline size_t perm_t4_smem_bytes(const PermDesc& d) {
  return (size_t(1) << d.tbits) + 8 * (size_t(1) << (d.tbits - d.a_bits)) + 8 * (size_t(1) << (d.tbits - d.b_bits)) +
         4 * (size_t(1) << (d.tbits - 6)) + 16;
}

inline void launch_permute_t4(tsb_ctx* ctx, const PermT4Desc& d, const void* src, void* dst, cudaStream_t stream) {
  const size_t smem = perm_t4_smem_bytes(d.p);
  const int grid = (int)std::min<int64_t>(d.p.n_tiles, (int64_t)ctx->sm_count * 8);
  TSB_LAUNCH(k_permute_t4, dim3(grid), dim3(kPermThreads), smem, stream, d, (const uint8_t*)src, (uint8_t*)dst);
  ctx->note("k_permute_t4");
  TSB_CUDA(cudaGetLastError());
}

inline size_t perm_smem_bytes(const PermDesc& d) {
  return (size_t(1) << d.tbits) + 8 * (size_t(1) << (d.tbits - d.a_bits)) + 8 * (size_t(1) << (d.tbits - d.b_bits)) +
         2 * (size_t(1) << (d.tbits - 4)) + 16;
}

inline void launch_permute_fast(tsb_ctx* ctx, const PermDesc& d, int G, const void* src, void* dst, cudaStream_t stream) {
  const size_t smem = perm_smem_bytes(d);
  const int grid = (int)std::min<int64_t>(d.n_tiles, (int64_t)ctx->sm_count * 8);
  const uint8_t* s = (const uint8_t*)src;
  uint8_t* t = (uint8_t*)dst;
  switch (G) {
    case 1: TSB_LAUNCH(k_permute<1>, dim3(grid), dim3(kPermThreads), smem, stream, d, s, t); break;
    case 2: TSB_LAUNCH(k_permute<2>, dim3(grid), dim3(kPermThreads), smem, stream, d, s, t); break;
    case 4: TSB_LAUNCH(k_permute<4>, dim3(grid), dim3(kPermThreads), smem, stream, d, s, t); break;
    case 8: TSB_LAUNCH(k_permute<8>, dim3(grid), dim3(kPermThreads), smem, stream, d, s, t); break;
    default: TSB_LAUNCH(k_permute<16>, dim3(grid), dim3(kPermThreads), smem, stream, d, s, t); break;
  }
  ctx->note(G == 1 ? "k_permute<1>" : G == 2 ? "k_permute<2>" : G == 4 ? "k_permute<4>" : G == 8 ? "k_permute<8>" : "k_permute<16>");
  TSB_CUDA(cudaGetLastError());
}

// General bit-permutation (or gather with holes): dst[dst_bits] = src[src_bits], where every dst label
// must occur in src_pos (label -> source ELEMENT address bit).  Chooses fast kernel / memcpy / generic.
inline void permute_bits(tsb_ctx* ctx, const void* src, const BitLabels& src_bits, void* dst,
                         const BitLabels& dst_bits, int elem_size, cudaStream_t stream) {
  TSB_CHECK(src_bits.size() == dst_bits.size(), TSB_ERR_INVALID, "permute: rank mismatch");
  const int n = (int)src_bits.size();
  if (src_bits == dst_bits) {
    if (src != dst)
      TSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)elem_size << n, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  PermDesc pd;
  PermT4Desc t4;
  bool use_t4 = false;
  int G = 1;
  if (aligned && !has_dups(src_bits) && make_perm_desc(src_bits, dst_bits, ilog2(elem_size), pd, G, &t4, &use_t4)) {
    static const int t4_mode = [] { const char* e = getenv("TSB_PERMUTE_T4"); return e ? atoi(e) : 1; }();  // tuning knob
    if (use_t4 && t4_mode) launch_permute_t4(ctx, t4, src, dst, stream);
    else launch_permute_fast(ctx, pd, G, src, dst, stream);
    return;
  }
  auto sp = positions(src_bits);
  std::vector<LDim> dims;
  for (int i = 0; i < n; i++) {
    LDim d{};
    d.label = i; d.ext = 2; d.in_l = d.in_o = true;
    d.so = int64_t(1) << (n - 1 - i);
    d.sl = int64_t(1) << sp.at(dst_bits[i]);
    dims.push_back(d);
  }
  const int dt = elem_size == 1 ? TSB_U8 : elem_size == 4 ? TSB_I32 : TSB_I64;
  TSB_CHECK(elem_size == 1 || elem_size == 4 || elem_size == 8, TSB_ERR_UNSUPPORTED, "element size %d", elem_size);
  GenDesc g = make_gen_desc(dims, dt, dt, dt, false, false, false);
  launch_generic(ctx, g, src, nullptr, dst, stream);
}

// Summed wires handled inside one merge launch; more are split off into a temporary leading axis of the
// result and summed by further passes, so that the work stays parallel over the output (split-K).
constexpr int kSplitKBits = 8;
// Arithmetic intensity (ops per algorithmic byte) above which a step goes to the tensor cores: measured on
// B200 the CUDA-core merge kernel needs about 0.1 ms per summed-wire combination and 2^30 outputs, the
// transposes around the GEMM about 0.8 ms per 2^30 outputs, so K >= 32 with wide M, N is the break-even.
constexpr double kGemmMinIntensity = 48.0;

// ---- merge lowering -------------------------------------------------------------------------
struct MergeArgs {
  MergeDesc d;
  bool ok = false;
};

// lhs_bits / rhs_bits / out_bits: bit labels MSB first.  Labels of lhs/rhs not in out are summed.
// Returns ok=false if the monotone-merge preconditions do not hold.
// allow_broadcast: out labels held by neither operand repeat the value (used to broadcast a packed operand).
// vbits: output address bits covered by one thread's vector (4 = 16 bytes of uint8 entries, 7 = 16 bytes of
// bit-packed entries); nibL / nibR are vbits-wide lane masks.
inline MergeArgs make_merge_desc(const BitLabels& lhs_bits, const BitLabels* rhs_bits, const BitLabels& out_bits,
                                 bool accumulate, int vbits = 4, bool allow_broadcast = false) {
  MergeArgs m;
  const int R = (int)out_bits.size();
  if (R < vbits || R > 62) return m;
  if (has_dups(lhs_bits) || has_dups(out_bits) || (rhs_bits && has_dups(*rhs_bits))) return m;
  auto lp = positions(lhs_bits);
  std::map<int64_t, int> rp;
  if (rhs_bits) rp = positions(*rhs_bits);
  auto op = positions(out_bits);
  // every out label must come from an operand
  if (!allow_broadcast)
    for (auto& kv : op)
      if (!lp.count(kv.first) && !rp.count(kv.first)) return m;
  // monotone: operand positions strictly increase with out position
  auto monotone = [&](const std::map<int64_t, int>& pp, int& nib, int& nlow) {
    int last = -1;
    nib = 0; nlow = 0;
    for (int j = 0; j < R; j++) {
      const int64_t lab = out_bits[R - 1 - j];
      auto it = pp.find(lab);
      if (it == pp.end()) continue;
      if (it->second <= last) return false;
      last = it->second;
      if (j < vbits) {
        if (it->second != nlow) return false;  // must be the operand's lowest address bits
        nib |= 1 << j;
        nlow++;
      }
    }
    return true;
  };
  int nibL = 0, nibR = 0, nl = 0, nr = 0;
  if (!monotone(lp, nibL, nl)) return m;
  if (rhs_bits && !monotone(rp, nibR, nr)) return m;
  MergeDesc& d = m.d;
  d = MergeDesc{};
  d.R = R;
  d.nibL = nibL; d.nibR = nibR;
  d.has_rhs = rhs_bits ? 1 : 0;
  d.accumulate = accumulate ? 1 : 0;
  const int vb = R - vbits;
  d.tile_bits = std::min(vb, std::min(kMergeTileBits, std::max(vbits == 4 ? 7 : 8, vb - 10)));
  d.n_vec = int64_t(1) << vb;
  for (int j = vbits; j < R; j++) {
    const int64_t lab = out_bits[R - 1 - j];
    const bool lo = (j - vbits) < d.tile_bits;
    const int in = lo ? j - vbits : j - vbits - d.tile_bits;
    auto il = lp.find(lab);
    if (il != lp.end()) runs_add(lo ? d.loL : d.hiL, in, il->second);
    auto ir = rp.find(lab);
    if (ir != rp.end()) runs_add(lo ? d.loR : d.hiR, in, ir->second);
  }
  // summed labels: everything in lhs/rhs that is not in out (lone labels get stride 0 on the other side)
  std::vector<int64_t> ks;
  for (auto& kv : lp) if (!op.count(kv.first)) ks.push_back(kv.first);
  for (auto& kv : rp) if (!op.count(kv.first) && !lp.count(kv.first)) ks.push_back(kv.first);
  if ((int)ks.size() > kMergeMaxK) return m;
  // iterate the lowest-address summed label fastest
  std::sort(ks.begin(), ks.end(), [&](int64_t a, int64_t b) {
    const int pa = lp.count(a) ? lp[a] : 64 + rp[a], pb = lp.count(b) ? lp[b] : 64 + rp[b];
    return pa < pb;
  });
  d.kbits = (int)ks.size();
  for (int q = 0; q < d.kbits; q++) {
    d.ksL[q] = lp.count(ks[q]) ? int64_t(1) << lp[ks[q]] : 0;
    d.ksR[q] = rp.count(ks[q]) ? int64_t(1) << rp[ks[q]] : 0;
  }
  m.ok = true;
  return m;
}

// Fused n-ary bit-packed merge: operands[i] / out are bit labels MSB first; every operand must hold the seven
// lowest output labels (plans store small operands broadcast), at most kBitsMaxKB summed labels.
struct BitsNArgs {
  BitsNDesc d;
  bool ok = false;
};

inline BitsNArgs make_bits_n_desc(const std::vector<BitLabels>& operands, const BitLabels& out_bits) {
  BitsNArgs m;
  const int R = (int)out_bits.size(), n = (int)operands.size();
  if (R < 7 || R > 62 || n < 2 || n > kBitsMaxOps || has_dups(out_bits)) return m;
  auto op = positions(out_bits);
  std::vector<std::map<int64_t, int>> lp(n);
  for (int o = 0; o < n; o++) {
    if (has_dups(operands[o])) return m;
    lp[o] = positions(operands[o]);
  }
  for (auto& kv : op) {
    bool found = false;
    for (int o = 0; o < n; o++) found = found || lp[o].count(kv.first);
    if (!found) return m;
  }
  std::vector<int64_t> ks;
  for (int o = 0; o < n; o++)
    for (auto& kv : lp[o])
      if (!op.count(kv.first) && std::find(ks.begin(), ks.end(), kv.first) == ks.end()) ks.push_back(kv.first);
  if ((int)ks.size() > kBitsMaxKB) return m;
  BitsNDesc& d = m.d;
  d = BitsNDesc{};
  d.R = R;
  d.n_ops = n;
  d.kbits = (int)ks.size();
  const int vb = R - 7;
  d.tile_bits = std::min(vb, std::min(kMergeTileBits, std::max(8, vb - 10)));
  d.n_vec = int64_t(1) << vb;
  for (int o = 0; o < n; o++) {
    BitsOperand& bo = d.op[o];
    int last = -1, nlow = 0, nib = 0;
    for (int j = 0; j < R; j++) {
      auto it = lp[o].find(out_bits[R - 1 - j]);
      if (it == lp[o].end()) continue;
      if (it->second <= last) return m;
      last = it->second;
      if (j < 7) {
        if (it->second != nlow) return m;
        nib |= 1 << j;
        nlow++;
      } else {
        const bool lo = (j - 7) < d.tile_bits;
        runs_add(lo ? bo.lo : bo.hi, lo ? j - 7 : j - 7 - d.tile_bits, it->second);
      }
    }
    if (nib != 127) return m;               // every operand holds the seven lane bits: one 16-byte load per fetch
    if (operands[o].size() > 38) return m;  // 16-byte-unit offsets are 32-bit
    bo.hi2 = nib >> 5;
    for (int q = 0; q < d.kbits; q++) bo.ks[q] = lp[o].count(ks[q]) ? int64_t(1) << lp[o][ks[q]] : 0;
  }
  m.ok = true;
  return m;
}

// Factored fused merge (k_merge_bits_f.cuh).  Same preconditions as make_bits_n_desc with at most kBitsFMaxKB
// summed labels; additionally sorts the operand slots by the summed labels they hold and picks up to
// `max_pair` output labels (above the lane bits) that as few operand fetches as possible depend on.
struct BitsFArgs {
  BitsFDesc d;
  bool ok = false;
  long wf_chosen = 0, wf_default = 0;  // modelled L1 wavefronts per work item: chosen lane wires / output bits 8..11
};

inline bool bits_f_sorted_host(int masks, int kb, int nops) {
  for (int o = 0; o + 1 < nops; o++)
    if (((masks >> (o * kb)) & ((1 << kb) - 1)) < ((masks >> ((o + 1) * kb)) & ((1 << kb) - 1))) return false;
  return true;
}

inline BitsFArgs make_bits_f_desc(const std::vector<BitLabels>& operands, const BitLabels& out_bits, int max_pair,
                                  int n_top_excluded = 0, bool choose_lanes = false, int choose_tile = 0) {
  BitsFArgs m;
  const int R = (int)out_bits.size(), n = (int)operands.size();
  if (R < 7 || R > 62 || n < 2 || n > kBitsMaxOps || has_dups(out_bits)) return m;
  auto op = positions(out_bits);
  std::vector<std::map<int64_t, int>> lp(n);
  for (int o = 0; o < n; o++) {
    if (has_dups(operands[o]) || operands[o].size() > 38) return m;
    lp[o] = positions(operands[o]);
  }
  for (auto& kv : op) {
    bool found = false;
    for (int o = 0; o < n; o++) found = found || lp[o].count(kv.first);
    if (!found) return m;
  }
  std::vector<int64_t> ks;
  for (int o = 0; o < n; o++)
    for (auto& kv : lp[o])
      if (!op.count(kv.first) && std::find(ks.begin(), ks.end(), kv.first) == ks.end()) ks.push_back(kv.first);
  const int KB = (int)ks.size();
  if (KB > kBitsFMaxKB) return m;
  // monotone order and the seven lane bits
  for (int o = 0; o < n; o++) {
    int last = -1, nlow = 0;
    for (int j = 0; j < R; j++) {
      auto it = lp[o].find(out_bits[R - 1 - j]);
      if (it == lp[o].end()) { if (j < 7) return m; continue; }
      if (it->second <= last) return m;
      last = it->second;
      if (j < 7 && it->second != nlow++) return m;
    }
  }
  // slots sorted by summed-label mask, descending
  std::vector<int> kmask(n, 0), slot(n);
  for (int o = 0; o < n; o++) {
    for (int q = 0; q < KB; q++) if (lp[o].count(ks[q])) kmask[o] |= 1 << q;
    slot[o] = o;
  }
  // pair labels: output positions >= 8 (vector bit >= 1 keeps the two 16-byte stores of a lane pair in one sector)
  const int vb = R - 7;
  const int min_env = [] { const char* e = getenv("TSB_BITSF_MIN_ITEMS_LOG2"); return e ? atoi(e) : -1; }();   // (tests: pair wires on small steps)
  const int min_items = min_env >= 0 ? min_env : 8;
  // Private streaming (max_pair >= 2, groups that have the kernels): two or three output wires that ONLY one operand
  // holds, that operand having the largest summed-wire mask.  It goes to slot 0 and is streamed over the pair-wire
  // combinations while the other operands are fetched once per work item.  Two-sided: a second operand with one or two
  // private wires of its own goes to slot 1 (if the slots stay sorted by mask); its vectors for the combinations of
  // those wires are folded in with the rest, and every vector streamed from slot 0 yields 2 / 4 results.  The highest
  // such wires are taken (the low output wires keep indexing lanes and tiles); the leading (sliced) wires stay index
  // bits (sub-range launches); at least 2^18 work items are kept (~7 tiles per SM at the smallest tile).
  int priv = -1, priv2 = -1;
  std::vector<int> pair_s, pair_t;
  auto private_wires = [&](int o, int cap) {
    std::vector<int> mine;
    for (int j = R - n_top_excluded - 1; j >= 8 && (int)mine.size() < cap; j--) {
      bool only = lp[o].count(out_bits[R - 1 - j]) > 0;
      for (int o2 = 0; o2 < n && only; o2++) only = o2 == o || !lp[o2].count(out_bits[R - 1 - j]);
      if (only) mine.push_back(j);
    }
    return mine;
  };
  if (max_pair >= 2 && (n == 3 || n == 4) && (KB == 1 || KB == 2)) {
    const int budget = vb - (min_env >= 0 ? min_env : 18);   // pair wires this step can afford
    const int top = *std::max_element(kmask.begin(), kmask.end());
    for (int o = 0; o < n; o++) {
      if (kmask[o] != top) continue;
      std::vector<int> mine = private_wires(o, std::min(std::min(max_pair, kBitsFMaxPair), budget));
      if (mine.size() >= 2 && mine.size() > pair_s.size()) { pair_s = mine; priv = o; }
    }
    const int two_env = [] { const char* e = getenv("TSB_BITSF_TWO_SIDED"); return e ? atoi(e) : 1; }();   // A/B knob: wires of slot 1 (measured: 1 beats 2 on the n=50 bench)
    if (priv >= 0 && KB == 2 && two_env > 0) {
      int second = -1;   // largest mask among the others
      for (int o = 0; o < n; o++) if (o != priv) second = std::max(second, kmask[o]);
      std::vector<int> cand_t;
      for (int o = 0; o < n; o++) {
        if (o == priv || kmask[o] != second) continue;
        std::vector<int> mine = private_wires(o, std::min(two_env, 2));
        if (!mine.empty() && mine.size() > cand_t.size()) { cand_t = mine; priv2 = o; }
      }
      if (priv2 >= 0) {
        // split of the budget between the two sides: fewest operand fetches per output vector
        double rest = 0, best = -1;
        for (int o = 0; o < n; o++) if (o != priv && o != priv2) rest += 1 << __builtin_popcount((unsigned)kmask[o]);
        const double f0 = 1 << __builtin_popcount((unsigned)kmask[priv]), f1 = 1 << __builtin_popcount((unsigned)kmask[priv2]);
        int ba = (int)pair_s.size(), bb = 0;
        for (int a = 2; a <= (int)pair_s.size(); a++)
          for (int b = 0; b <= (int)cand_t.size() && a + b <= budget; b++) {
            const double cost = (f0 * (1 << a) + f1 * (1 << b) + rest) / (double)(1 << (a + b));
            if (best < 0 || cost < best) { best = cost; ba = a; bb = b; }
          }
        pair_s.resize(ba);
        pair_t.assign(cand_t.begin(), cand_t.begin() + bb);
        if (bb == 0) priv2 = -1;
      }
    }
  }
  std::stable_sort(slot.begin(), slot.end(), [&](int a, int b) {
    if ((a == priv) != (b == priv)) return a == priv;
    if ((a == priv2) != (b == priv2)) return a == priv2;
    return kmask[a] > kmask[b];
  });
  std::vector<int> pair_pos;
  if (priv < 0) {
    // one pair wire, chosen among the output wires by the operand fetches that do not depend on it
    std::vector<int> loads(n);
    for (int o = 0; o < n; o++) loads[o] = 1 << __builtin_popcount((unsigned)kmask[o]);
    if (max_pair >= 1 && vb > min_items) {
      int best = -1, best_score = 0;
      for (int j = 8; j < R - n_top_excluded; j++) {  // the leading (sliced) wires stay index bits: sub-range launches
        int score = 0;
        for (int o = 0; o < n; o++)
          if (!lp[o].count(out_bits[R - 1 - j])) score += loads[o];
        if (score > best_score) { best_score = score; best = j; }
      }
      if (best >= 0) pair_pos.push_back(best);
    }
  } else {
    std::sort(pair_s.begin(), pair_s.end());
    std::sort(pair_t.begin(), pair_t.end());
    pair_pos = pair_s;
    pair_pos.insert(pair_pos.end(), pair_t.begin(), pair_t.end());
  }
  const int NP = (int)pair_pos.size();   // all pair wires; slot 0's come first
  BitsFDesc& d = m.d;
  d = BitsFDesc{};
  d.R = R; d.kbits = KB; d.n_ops = n;
  d.n_pair = priv < 0 ? NP : (int)pair_s.size();
  d.n_pair2 = (int)pair_t.size();
  const int vr = vb - NP;  // reduced vector-index bits
  static const int tile_env = [] { const char* e = getenv("TSB_BITSF_TILE_BITS"); return e ? atoi(e) : kMergeTileBits; }();  // A/B knob
  // (private streaming does 4 ... 32 vectors per work item: smaller tiles keep at least ~7 tiles per SM)
  d.tile_bits = std::min(vr, std::min(std::min(kMergeTileBits, std::max(tile_env, 6)), std::max(8, vr - (NP >= 2 ? 11 : 10))));
  d.n_vec = int64_t(1) << vr;
  for (int g = 0; g < 8; g++) {
    uint32_t w = 0;
    for (int q = 0; q < d.n_pair; q++) if (g >> q & 1) w += uint32_t(1) << (pair_pos[q] - 7);
    d.ow[g] = w;
  }
  for (int h = 0; h < 4; h++) {
    uint32_t w = 0;
    for (int q = 0; q < d.n_pair2; q++) if (h >> q & 1) w += uint32_t(1) << (pair_pos[d.n_pair + q] - 7);
    d.ow2[h] = w;
  }
  for (int s = 0; s < n; s++) {
    const int o = slot[s];
    BitsFOperand& bo = d.op[s];
    d.perm[s] = o;
    bo.kmask = kmask[o];
    d.masks |= kmask[o] << (s * KB);
    for (int kk = 0; kk < 4; kk++) {
      int64_t a = 0;
      for (int q = 0; q < KB; q++) if ((kk >> q & 1) && lp[o].count(ks[q])) a += int64_t(1) << lp[o][ks[q]];
      bo.kw[kk] = (uint32_t)(a >> 7);
    }
    std::vector<int> mine;   // the pair wires this operand holds, in pair_pos order
    for (int q = 0; q < NP; q++)
      if (lp[o].count(out_bits[R - 1 - pair_pos[q]])) { bo.held |= 1 << q; mine.push_back(pair_pos[q]); }
    for (int g = 0; g < 8; g++) {
      int64_t a = 0;
      for (int q = 0; q < (int)mine.size() && q < 3; q++)
        if (g >> q & 1) a += int64_t(1) << lp[o][out_bits[R - 1 - mine[q]]];
      bo.pw[g] = (uint32_t)(a >> 7);
    }
  }
  if (priv >= 0 && !bits_f_sorted_host(d.masks, KB, n)) {
    // slot 1's mask is not the second largest after all: one-sided (cannot happen with the choice above)
    return make_bits_f_desc(operands, out_bits, 1, n_top_excluded, choose_lanes, choose_tile);
  }
  // Which output wires index the 32 lanes of a warp.  A warp-wide 16-byte-per-lane access costs one L1 wavefront
  // per distinct 128-byte line, and these steps are bound by exactly that.  Lane bit 0 stays the lowest vector bit
  // (two lanes fill one 32-byte sector of the result); the other four are chosen, by exhaustive search over the
  // eligible output wires, to minimise wavefronts per work item: an operand that does not hold a lane wire, or
  // holds it among its own address bits 7..9, is fetched from fewer lines.
  std::vector<int> lanes;
  if (choose_lanes && vr >= 6) {
    std::vector<int> cand;
    for (int j = 8; j < R - n_top_excluded; j++)
      if (std::find(pair_pos.begin(), pair_pos.end(), j) == pair_pos.end()) cand.push_back(j);
    const int C = (int)cand.size();
    if (C >= 4) {
      // per operand and candidate: held?, and whether it lands within the operand's first 128 bytes (position <= 9)
      std::vector<std::vector<int>> cls(n, std::vector<int>(C, 0));  // 0: not held, 1: held in-line, 2: held, new line
      std::vector<int> fetch(n);
      for (int o = 0; o < n; o++) {
        fetch[o] = 1 << (__builtin_popcount((unsigned)kmask[o]));
        for (int q = 0; q < NP; q++) if (lp[o].count(out_bits[R - 1 - pair_pos[q]])) fetch[o] *= 2;
        for (int c = 0; c < C; c++) {
          auto it = lp[o].find(out_bits[R - 1 - cand[c]]);
          if (it != lp[o].end()) cls[o][c] = it->second <= 9 ? 1 : 2;
        }
      }
      long best_cost = -1;
      int bi[4] = {0, 1, 2, 3};
      for (int a = 0; a < C; a++)
        for (int b = a + 1; b < C; b++)
          for (int c = b + 1; c < C; c++)
            for (int e = c + 1; e < C; e++) {
              const int pick[4] = {a, b, c, e};
              // result lines per warp store: contiguous run of lane wires from output bit 7 upwards
              int run = 1;
              for (int q = 0; q < 4 && cand[pick[q]] == 7 + run; q++) run++;
              long cost = (long)(1 << NP) * (run >= 3 ? 4 : run == 2 ? 8 : 16);
              for (int o = 0; o < n; o++) {
                int far = 0;
                for (int q = 0; q < 4; q++) far += cls[o][pick[q]] == 2;
                cost += (long)fetch[o] << far;
              }
              if (a == 0 && b == 1 && c == 2 && e == 3) m.wf_default = cost;
              if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                for (int q = 0; q < 4; q++) bi[q] = pick[q];
              }
            }
      m.wf_chosen = best_cost;
      for (int q = 0; q < 4; q++) lanes.push_back(cand[bi[q]]);
    }
  }
  // Which output wires index the vectors of one CTA tile.  The operands of these steps are small next to the result
  // (each lacks many output wires): with the tile spanned by wires that FEW operand fetches depend on, the part of
  // every operand a tile touches shrinks to a few KB and is re-read from L1 instead of L2.  Greedy: starting from
  // output bit 7, add the wire whose holders have the smallest footprint so far (ties: the lower bit).
  if (choose_tile && lanes.empty() && vr > d.tile_bits) {
    std::vector<double> fp(n);
    for (int o = 0; o < n; o++) {
      fp[o] = (double)(1 << __builtin_popcount((unsigned)kmask[o]));
      for (int q = 0; q < NP; q++) if (lp[o].count(out_bits[R - 1 - pair_pos[q]])) fp[o] *= 2;
      if (lp[o].count(out_bits[R - 1 - 7])) fp[o] *= 2;
    }
    std::vector<int> cand;
    for (int j = 8; j < R - n_top_excluded; j++)
      if (std::find(pair_pos.begin(), pair_pos.end(), j) == pair_pos.end()) cand.push_back(j);
    if (choose_tile == 2) {
      // the 32 lanes of a warp keep the lowest output wires (coalesced 16-byte fetches and stores); only the
      // remaining tile wires are chosen by footprint
      while ((int)lanes.size() < 4 && !cand.empty()) {
        const int j = cand.front();
        cand.erase(cand.begin());
        lanes.push_back(j);
        for (int o = 0; o < n; o++) if (lp[o].count(out_bits[R - 1 - j])) fp[o] *= 2;
      }
    }
    while ((int)lanes.size() < d.tile_bits - 1 && !cand.empty()) {
      int best = -1;
      double best_cost = 0;
      for (int c = 0; c < (int)cand.size(); c++) {
        double cost = 0;
        for (int o = 0; o < n; o++) if (lp[o].count(out_bits[R - 1 - cand[c]])) cost += fp[o];
        if (best < 0 || cost < best_cost) { best = c; best_cost = cost; }
      }
      const int j = cand[best];
      cand.erase(cand.begin() + best);
      lanes.push_back(j);
      for (int o = 0; o < n; o++) if (lp[o].count(out_bits[R - 1 - j])) fp[o] *= 2;
    }
    double total = 0;
    for (int o = 0; o < n; o++) total += fp[o];
    m.wf_chosen = (long)total;   // 16-byte vectors of all operands one tile touches
  }
  std::vector<int> index_bits;  // reduced index bit -> output bit
  index_bits.push_back(7);
  for (int j : lanes) index_bits.push_back(j);
  for (int j = 8; j < R; j++)
    if (std::find(pair_pos.begin(), pair_pos.end(), j) == pair_pos.end() &&
        std::find(lanes.begin(), lanes.end(), j) == lanes.end())
      index_bits.push_back(j);
  for (int in = 0; in < (int)index_bits.size(); in++) {
    const int j = index_bits[in];
    const bool lo = in < d.tile_bits;
    const int ii = lo ? in : in - d.tile_bits;
    runs_add(lo ? d.out_lo : d.out_hi, ii, j - 7);
    for (int s = 0; s < n; s++) {
      auto it = lp[slot[s]].find(out_bits[R - 1 - j]);
      if (it != lp[slot[s]].end()) runs_add(lo ? d.op[s].lo : d.op[s].hi, ii, it->second);
    }
  }
  // deposit_warp handles one run per lane
  if (d.out_hi.n > 32) return m;
  for (int s = 0; s < n; s++)
    if (d.op[s].hi.n > 32) return m;
  m.ok = true;
  return m;
}

void launch_merge_bits_f(tsb_ctx* ctx, const BitsFDesc& d, const void* const* ptrs, const uint64_t* bitoffs, void* o,
                         cudaStream_t s, const uint32_t* tables = nullptr);
std::vector<uint32_t> bits_f_tables(const BitsFDesc& d);

void launch_merge_bits_n(tsb_ctx* ctx, const BitsNDesc& d, const void* const* ptrs, const uint64_t* bitoffs, void* o,
                         cudaStream_t s);

#ifdef TSB_MERGE_GENERIC_IMPL
template <typename TL, typename TR, typename TO>
inline void launch_merge_t(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, bool clamp,
                           cudaStream_t stream) {
  const int64_t n_tiles = d.n_vec >> d.tile_bits;
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 16);
  auto kc = k_merge<TL, TR, TO, 1>;
  auto kn = k_merge<TL, TR, TO, 0>;
  if (clamp)
    TSB_LAUNCH(kc, dim3(grid), dim3(kMergeThreads), 0, stream, d, (const TL*)l, (const TR*)r, (TO*)o);
  else
    TSB_LAUNCH(kn, dim3(grid), dim3(kMergeThreads), 0, stream, d, (const TL*)l, (const TR*)r, (TO*)o);
  ctx->note("k_merge");
  TSB_CUDA(cudaGetLastError());
}

template <typename TL, typename TR>
inline void launch_merge_o(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, int dout,
                           bool clamp, cudaStream_t s) {
  if (dout == TSB_U8) launch_merge_t<TL, TR, uint8_t>(ctx, d, l, r, o, clamp, s);
  else if (dout == TSB_I32) launch_merge_t<TL, TR, int32_t>(ctx, d, l, r, o, clamp, s);
  else launch_merge_t<TL, TR, int64_t>(ctx, d, l, r, o, clamp, s);
}

template <typename TL>
inline void launch_merge_r(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, int dr, int dout,
                           bool clamp, cudaStream_t s) {
  if (dr == TSB_U8) launch_merge_o<TL, uint8_t>(ctx, d, l, r, o, dout, clamp, s);
  else if (dr == TSB_I32) launch_merge_o<TL, int32_t>(ctx, d, l, r, o, dout, clamp, s);
  else launch_merge_o<TL, int64_t>(ctx, d, l, r, o, dout, clamp, s);
}
#endif

bool merge_sat_supported(const MergeDesc& d);
void launch_merge_sat(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t stream);
void launch_merge_generic(tsb_ctx* ctx, const MergeDesc& d, const void* l, int dl, const void* r, int dr, void* o,
                          int dout, bool clamp, cudaStream_t s);

template <typename TL, typename TR>
void launch_merge_cnt_t(tsb_ctx* ctx, const MergeDesc& d, const void* l, const void* r, void* o, cudaStream_t stream);
template <> void launch_merge_cnt_t<uint8_t, uint8_t>(tsb_ctx*, const MergeDesc&, const void*, const void*, void*, cudaStream_t);
template <> void launch_merge_cnt_t<uint8_t, int32_t>(tsb_ctx*, const MergeDesc&, const void*, const void*, void*, cudaStream_t);
template <> void launch_merge_cnt_t<int32_t, uint8_t>(tsb_ctx*, const MergeDesc&, const void*, const void*, void*, cudaStream_t);
template <> void launch_merge_cnt_t<int32_t, int32_t>(tsb_ctx*, const MergeDesc&, const void*, const void*, void*, cudaStream_t);

// bit-packed boolean relations (k_merge_bits.cuh); d from make_merge_desc(..., vbits = 7)
void launch_merge_bits(tsb_ctx* ctx, const MergeDesc& d, const void* l, uint64_t bitoff_l, const void* r,
                       uint64_t bitoff_r, void* o, cudaStream_t s);
void launch_pack_bits(tsb_ctx* ctx, const void* src_u8, void* dst_bits, int64_t n_entries, cudaStream_t s);
void launch_unpack_bits(tsb_ctx* ctx, const void* src_bits, void* dst_u8, int64_t n_entries, bool accumulate,
                        cudaStream_t s);
// dst bit i = src entry (deposit(hi, i >> 5) + deposit(lo, i & 31)); src is uint8 entries or bit-packed
void launch_pack_map(tsb_ctx* ctx, const void* src, bool src_packed, void* dst_bits, int64_t n_dst_entries,
                     const BitRuns& lo, const BitRuns& hi, cudaStream_t s);
inline size_t packed_bytes(int64_t n_entries) { return (size_t)std::max<int64_t>(16, (n_entries + 7) / 8); }

bool merge_wide_supported(const MergeDesc& d, const void* l, int dl, const void* r, int dr, const void* o);
void launch_merge_wide(tsb_ctx* ctx, const MergeDesc& d, const void* l, int dl, const void* r, int dr, void* o,
                       cudaStream_t s);

inline void launch_merge(tsb_ctx* ctx, const MergeDesc& d, const void* l, int dl, const void* r, int dr, void* o,
                         int dout, bool clamp, cudaStream_t s) {
  if (clamp && dl == TSB_U8 && dr == TSB_U8 && dout == TSB_U8 && merge_sat_supported(d)) {
    launch_merge_sat(ctx, d, l, r, o, s);
    return;
  }
  if (!clamp && dout == TSB_I32 && dl != TSB_I64 && dr != TSB_I64 && merge_sat_supported(d)) {
    // counting semiring, int32 results: compile-time specialised kernel
    if (dl == TSB_U8 && dr == TSB_U8) launch_merge_cnt_t<uint8_t, uint8_t>(ctx, d, l, r, o, s);
    else if (dl == TSB_U8) launch_merge_cnt_t<uint8_t, int32_t>(ctx, d, l, r, o, s);
    else if (dr == TSB_U8) launch_merge_cnt_t<int32_t, uint8_t>(ctx, d, l, r, o, s);
    else launch_merge_cnt_t<int32_t, int32_t>(ctx, d, l, r, o, s);
    return;
  }
  if (!clamp && dout == TSB_I64 && merge_wide_supported(d, l, dl, r, dr, o)) {
    launch_merge_wide(ctx, d, l, dl, r, dr, o, s);  // counting semiring, int64 results
    return;
  }
  launch_merge_generic(ctx, d, l, dl, r, dr, o, dout, clamp, s);
}

#ifdef TSB_MERGE_GENERIC_IMPL
void launch_merge_generic(tsb_ctx* ctx, const MergeDesc& d, const void* l, int dl, const void* r, int dr, void* o,
                          int dout, bool clamp, cudaStream_t s) {
  if (!d.has_rhs) dr = TSB_U8;
  if (dl == TSB_U8) launch_merge_r<uint8_t>(ctx, d, l, r, o, dr, dout, clamp, s);
  else if (dl == TSB_I32) launch_merge_r<int32_t>(ctx, d, l, r, o, dr, dout, clamp, s);
  else launch_merge_r<int64_t>(ctx, d, l, r, o, dr, dout, clamp, s);
}
#endif

}  // namespace tsb
