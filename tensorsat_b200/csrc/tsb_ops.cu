// tsb_ops.cu — C-ABI entry points for contexts, the arena and the single box operations.
#include <cmath>

#include "lower.h"

namespace tsb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

#define TSB_API_BEGIN try {
#define TSB_API_END                                          \
  return TSB_OK;                                             \
  }                                                          \
  catch (const ::tsb::Error& e) { return e.code; }           \
  catch (const std::bad_alloc&) {                            \
    ::tsb::set_error("host allocation failed");              \
    return TSB_ERR_OOM;                                      \
  }                                                          \
  catch (const std::exception& e) {                          \
    ::tsb::set_error("internal error: %s", e.what());        \
    return TSB_ERR_INVALID;                                  \
  }

static void check_ctx(tsb_ctx* ctx) { TSB_CHECK(ctx != nullptr, TSB_ERR_INVALID, "null context"); }
static void check_dtype(int dt) {
  TSB_CHECK(dt == TSB_U8 || dt == TSB_I32 || dt == TSB_I64, TSB_ERR_INVALID, "invalid dtype %d", dt);
}
static void check_mode(int mode) {
  TSB_CHECK(mode == TSB_SAT || mode == TSB_COUNT, TSB_ERR_INVALID, "invalid mode %d", mode);
}

// C-order element strides of a dense tensor
static std::vector<int64_t> c_strides(const int32_t* ext, int rank) {
  std::vector<int64_t> s(rank);
  int64_t acc = 1;
  for (int i = rank - 1; i >= 0; i--) {
    s[i] = acc;
    acc *= ext[i];
  }
  return s;
}

static int64_t numel(const int32_t* ext, int rank) {
  int64_t n = 1;
  for (int i = 0; i < rank; i++) n *= ext[i];
  return n;
}

#ifndef TSB_HOST_EMU
void gemm_u8_impl(tsb_ctx* ctx, const void* A, const void* B, void* C, int b_bits, int m_bits, int n_bits, int k_bits,
                  int out_dtype, int mode, cudaStream_t stream);
#endif

// Label classes of a pairwise step over bit labels, and whether the step is worth the tensor cores:
// arithmetic intensity above `min_intensity` op/B and a tile the UMMA shape can fill.
struct GemmShape {
  bool ok = false;
  bool swap = false;  // use rhs as the M side
  BitLabels B, M, N, K;
};

GemmShape gemm_shape(const BitLabels& lb, const BitLabels& rb, const BitLabels& ob, int out_elem_size,
                     double min_intensity, bool boolean_operands) {
  GemmShape g;
  std::set<int64_t> sl(lb.begin(), lb.end()), sr(rb.begin(), rb.end()), so(ob.begin(), ob.end());
  for (auto x : ob) {
    const bool inl = sl.count(x) != 0, inr = sr.count(x) != 0;
    if (inl && inr) g.B.push_back(x);
    else if (inl) g.M.push_back(x);
    else if (inr) g.N.push_back(x);
    else return g;
  }
  for (auto x : lb)
    if (!so.count(x)) {
      if (!sr.count(x)) return g;  // lone label: not a plain GEMM
      g.K.push_back(x);
    }
  for (auto x : rb)
    if (!so.count(x) && !sl.count(x)) return g;
  int m = (int)g.M.size(), n = (int)g.N.size();
  const int k = (int)g.K.size(), b = (int)g.B.size();
  if (m < n) { std::swap(g.M, g.N); std::swap(m, n); g.swap = true; }
  // s32 accumulation is exact for 0/1 operands up to K = 2^24, for general u8 operands up to K = 2^15
  if (m < 7 || n < 5 || k < 5 || k > (boolean_operands ? 24 : 15)) return g;
  const double ops = 2.0 * std::exp2((double)(b + m + n + k));
  const double bytes = std::exp2((double)(b + m + k)) + std::exp2((double)(b + n + k)) +
                       out_elem_size * std::exp2((double)(b + m + n));
  g.ok = ops / bytes >= min_intensity;
  return g;
}

// Generic-kernel descriptor of an einsum-like pairwise step (any extents, repeated / dropped labels).
GenDesc lower_generic_contract(int lhs_rank, const int32_t* lhs_wires, int rhs_rank, const int32_t* rhs_wires,
                               int out_rank, const int32_t* out_wires, const int32_t* wext, int dl, int dr, int dout,
                               bool clamp, bool accumulate, bool& out_dups) {
  std::vector<int32_t> le(lhs_rank), re(rhs_rank), oe(out_rank);
  for (int i = 0; i < lhs_rank; i++) le[i] = wext[lhs_wires[i]];
  for (int i = 0; i < rhs_rank; i++) re[i] = wext[rhs_wires[i]];
  for (int i = 0; i < out_rank; i++) oe[i] = wext[out_wires[i]];
  std::map<int32_t, LDim> dims;
  auto ls = c_strides(le.data(), lhs_rank), rs = c_strides(re.data(), rhs_rank), os = c_strides(oe.data(), out_rank);
  for (int i = 0; i < lhs_rank; i++) {
    LDim& d = dims[lhs_wires[i]];
    d.label = lhs_wires[i]; d.ext = le[i]; d.sl += ls[i]; d.in_l = true;
  }
  for (int i = 0; i < rhs_rank; i++) {
    LDim& d = dims[rhs_wires[i]];
    d.label = rhs_wires[i]; d.ext = re[i]; d.sr += rs[i]; d.in_r = true;
  }
  out_dups = false;
  for (int i = 0; i < out_rank; i++) {
    LDim& d = dims[out_wires[i]];
    if (d.in_o) out_dups = true;
    d.so += os[i]; d.in_o = true;
  }
  std::vector<LDim> v;
  for (auto& kv : dims) v.push_back(kv.second);
  return make_gen_desc(v, dl, dr, dout, clamp, true, accumulate);
}

// ---------------------------------------------------------------------------------------------
// contract2 (FinRel._contract2, tensorsat/lang/fin_rel.py:255-288)
// ---------------------------------------------------------------------------------------------
void contract2_impl(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires, int dl,
                    const void* rhs, int rhs_rank, const int32_t* rhs_wires, int dr, void* out, int out_rank,
                    const int32_t* out_wires, int dout, int n_wires, const int32_t* wext, int mode,
                    bool accumulate, bool force_generic, cudaStream_t stream) {
  TSB_CHECK(lhs_rank >= 0 && lhs_rank <= TSB_MAX_RANK && rhs_rank >= 0 && rhs_rank <= TSB_MAX_RANK &&
                out_rank >= 0 && out_rank <= TSB_MAX_RANK,
            TSB_ERR_INVALID, "rank out of range (max %d)", TSB_MAX_RANK);
  check_dtype(dl); check_dtype(dr); check_dtype(dout); check_mode(mode);
  auto chk = [&](const int32_t* w, int n, const char* what) {
    for (int i = 0; i < n; i++) {
      TSB_CHECK(w[i] >= 0 && w[i] < n_wires, TSB_ERR_INVALID, "%s wire %d out of range", what, w[i]);
      TSB_CHECK(wext[w[i]] >= 1, TSB_ERR_INVALID, "wire %d has extent %d", w[i], wext[w[i]]);
    }
  };
  chk(lhs_wires, lhs_rank, "lhs"); chk(rhs_wires, rhs_rank, "rhs"); chk(out_wires, out_rank, "out");
  std::vector<int32_t> le(lhs_rank), re(rhs_rank), oe(out_rank);
  for (int i = 0; i < lhs_rank; i++) le[i] = wext[lhs_wires[i]];
  for (int i = 0; i < rhs_rank; i++) re[i] = wext[rhs_wires[i]];
  for (int i = 0; i < out_rank; i++) oe[i] = wext[out_wires[i]];
  std::set<int32_t> in_l(lhs_wires, lhs_wires + lhs_rank), in_r(rhs_wires, rhs_wires + rhs_rank);
  for (int i = 0; i < out_rank; i++)
    TSB_CHECK(in_l.count(out_wires[i]) || in_r.count(out_wires[i]), TSB_ERR_INVALID,
              "Every output wire must appear in LHR or RHS.");
  const bool clamp = mode == TSB_SAT;
  const int64_t n_out = numel(oe.data(), out_rank);

  // ---- fast path: binary extents, monotone merge (with operand permutes when needed)
  const bool binary = !force_generic && all_pow2(le.data(), lhs_rank) && all_pow2(re.data(), rhs_rank) &&
                      ((((uintptr_t)lhs | (uintptr_t)rhs | (uintptr_t)out) & 15) == 0);
  bool fast = binary && n_out >= 256;
  if (binary) {
    BitLabels lb = bit_labels(lhs_wires, lhs_rank, wext), rb = bit_labels(rhs_wires, rhs_rank, wext),
              ob = bit_labels(out_wires, out_rank, wext);
    const bool dups = has_dups(lb) || has_dups(rb) || has_dups(ob);
    if (dups) fast = false;
#ifndef TSB_HOST_EMU
    // ---- tensor cores: steps above the ridge go through transpose -> int8 GEMM -> transpose
    if (fast && !accumulate && dl == TSB_U8 && dr == TSB_U8 && (clamp ? dout == TSB_U8 : dout != TSB_U8)) {
      GemmShape g = gemm_shape(lb, rb, ob, dtype_size(dout), kGemmMinIntensity, clamp);
      if (g.ok) {
        const void* pa = g.swap ? rhs : lhs;
        const void* pb = g.swap ? lhs : rhs;
        const BitLabels& ab = g.swap ? rb : lb;
        const BitLabels& bb = g.swap ? lb : rb;
        auto cat = [](const BitLabels& x, const BitLabels& y, const BitLabels& z) {
          BitLabels r = x; r.insert(r.end(), y.begin(), y.end()); r.insert(r.end(), z.begin(), z.end()); return r;
        };
        const BitLabels a_can = cat(g.B, g.M, g.K), b_can = cat(g.B, g.N, g.K), c_can = cat(g.B, g.M, g.N);
        std::unique_ptr<Temp> ta, tb, tc, tc2;
        if (ab != a_can) {
          ta.reset(new Temp(ctx, (size_t)1 << ab.size(), stream));
          permute_bits(ctx, pa, ab, ta->p, a_can, 1, stream);
          pa = ta->p;
        }
        if (bb != b_can) {
          tb.reset(new Temp(ctx, (size_t)1 << bb.size(), stream));
          permute_bits(ctx, pb, bb, tb->p, b_can, 1, stream);
          pb = tb->p;
        }
        const int gemm_dt = clamp ? TSB_U8 : TSB_I32;
        void* pc = out;
        if (c_can != ob || gemm_dt != dout) {
          tc.reset(new Temp(ctx, (size_t)dtype_size(gemm_dt) << ob.size(), stream));
          pc = tc->p;
        }
        gemm_u8_impl(ctx, pa, pb, pc, (int)g.B.size(), (int)g.M.size(), (int)g.N.size(), (int)g.K.size(), gemm_dt,
                     mode, stream);
        if (pc != out) {
          void* src = pc;
          if (gemm_dt != dout) {  // counting mode with an int64 result: widen after the transpose-free case
            if (c_can == ob) {
              TSB_LAUNCH(k_convert, dim3(grid_for(ctx, n_out, 256)), dim3(256), 0, stream, (const void*)src, gemm_dt, out,
                         dout, n_out, 0);
              ctx->note("k_convert");
              return;
            }
            tc2.reset(new Temp(ctx, (size_t)dtype_size(dout) << ob.size(), stream));
            TSB_LAUNCH(k_convert, dim3(grid_for(ctx, n_out, 256)), dim3(256), 0, stream, (const void*)src, gemm_dt,
                       tc2->p, dout, n_out, 0);
            ctx->note("k_convert");
            src = tc2->p;
          }
          permute_bits(ctx, src, c_can, out, ob, dtype_size(dout), stream);
        }
        return;
      }
    }
#endif
    if (!dups) {
      // ---- split-K: more than kSplitKBits summed labels -> keep the upper ones as leading result axes,
      // then sum them away pass by pass (each pass is a merge against the scalar 1)
      std::set<int64_t> oset(ob.begin(), ob.end()), lset(lb.begin(), lb.end()), rset(rb.begin(), rb.end());
      BitLabels ks;
      for (auto x : lb) if (!oset.count(x)) ks.push_back(x);
      for (auto x : rb) if (!oset.count(x) && !lset.count(x)) ks.push_back(x);
      if ((int)ks.size() > kSplitKBits) {
        BitLabels keep(ks.begin(), ks.begin() + ((int)ks.size() - kSplitKBits));  // most significant summed labels
        BitLabels tb = keep;
        tb.insert(tb.end(), ob.begin(), ob.end());
        if ((size_t)(clamp ? 1 : dtype_size(dout)) << std::min<size_t>(tb.size(), 40) <= ((size_t)1 << 34) && tb.size() <= 34) {
          const int tdt = clamp ? TSB_U8 : dout;
          Temp tmp(ctx, (size_t)dtype_size(tdt) << tb.size(), stream);
          // labels -> fresh wire ids for the recursive calls (one wire per bit label)
          std::map<int64_t, int32_t> id;
          auto wires_of = [&](const BitLabels& b) {
            std::vector<int32_t> w;
            for (auto x : b) { auto it = id.find(x); if (it == id.end()) it = id.emplace(x, (int32_t)id.size()).first; w.push_back(it->second); }
            return w;
          };
          std::vector<int32_t> wl = wires_of(lb), wr = wires_of(rb), wt = wires_of(tb), wo = wires_of(ob);
          std::vector<int32_t> ext(id.size(), 2);
          contract2_impl(ctx, lhs, (int)wl.size(), wl.data(), dl, rhs, (int)wr.size(), wr.data(), dr, tmp.p, (int)wt.size(),
                         wt.data(), tdt, (int)ext.size(), ext.data(), mode, false, false, stream);
          contract2_impl(ctx, tmp.p, (int)wt.size(), wt.data(), tdt, ctx->d_one, 0, nullptr, TSB_U8, out, (int)wo.size(),
                         wo.data(), dout, (int)ext.size(), ext.data(), mode, accumulate, false, stream);
          return;
        }
      }
    }
    if (fast) {
      MergeArgs m = make_merge_desc(lb, &rb, ob, accumulate);
      std::unique_ptr<Temp> tl, tr;
      const void* pl = lhs;
      const void* pr = rhs;
      if (!m.ok) {
        // bring each operand to [summed labels (own order)] [free labels in output order] when it is not
        // already mergeable as it stands
        std::set<int64_t> oset(ob.begin(), ob.end());
        auto canon = [&](const BitLabels& b) {
          BitLabels c;
          for (auto x : b) if (!oset.count(x)) c.push_back(x);
          std::set<int64_t> have(b.begin(), b.end());
          for (auto x : ob) if (have.count(x)) c.push_back(x);
          return c;
        };
        BitLabels lc = canon(lb), rc = canon(rb);
        // summed labels shared by both operands must not interleave differently: they are handled by
        // per-label strides, so any order works.
        if (lc != lb) {
          tl.reset(new Temp(ctx, (size_t)dtype_size(dl) << lb.size(), stream));
          permute_bits(ctx, lhs, lb, tl->p, lc, dtype_size(dl), stream);
          pl = tl->p; lb = lc;
        }
        if (rc != rb) {
          tr.reset(new Temp(ctx, (size_t)dtype_size(dr) << rb.size(), stream));
          permute_bits(ctx, rhs, rb, tr->p, rc, dtype_size(dr), stream);
          pr = tr->p; rb = rc;
        }
        m = make_merge_desc(lb, &rb, ob, accumulate);
      }
      if (m.ok) {
        launch_merge(ctx, m.d, pl, dl, pr, dr, out, dout, clamp, stream);
        return;
      }
    }
  }

  // ---- generic path
  bool out_dups = false;
  GenDesc g = lower_generic_contract(lhs_rank, lhs_wires, rhs_rank, rhs_wires, out_rank, out_wires, wext, dl, dr, dout,
                                     clamp, accumulate, out_dups);
  if (out_dups && !accumulate) TSB_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * dtype_size(dout), stream));
  launch_generic(ctx, g, lhs, rhs, out, stream);
}

// ---------------------------------------------------------------------------------------------
// rewire (FinRel._rewire / rewire_array, fin_rel.py:379-393, _utils/misc.py:104-123) and slicing
// ---------------------------------------------------------------------------------------------
// dst[j-th axis] <- src axis out_ports[j]; src axes with fixed_idx >= 0 are indexed, other absent axes summed.
void rewire_impl(tsb_ctx* ctx, const void* src, int src_rank, const int32_t* sext, const int32_t* fixed_idx,
                 int ds, void* dst, int out_rank, const int32_t* out_ports, int dd, bool clamp, bool accumulate,
                 cudaStream_t stream) {
  TSB_CHECK(src_rank >= 0 && src_rank <= TSB_MAX_RANK && out_rank >= 0 && out_rank <= TSB_MAX_RANK, TSB_ERR_INVALID,
            "rank out of range (max %d)", TSB_MAX_RANK);
  check_dtype(ds); check_dtype(dd);
  for (int i = 0; i < src_rank; i++) TSB_CHECK(sext[i] >= 1, TSB_ERR_INVALID, "axis %d has extent %d", i, sext[i]);
  auto ss = c_strides(sext, src_rank);
  std::vector<int32_t> oext(out_rank);
  for (int j = 0; j < out_rank; j++) {
    TSB_CHECK(out_ports[j] >= 0 && out_ports[j] < src_rank, TSB_ERR_INVALID, "Output ports must be valid for this box.");
    TSB_CHECK(!fixed_idx || fixed_idx[out_ports[j]] < 0, TSB_ERR_INVALID, "an indexed axis cannot be an output port");
    oext[j] = sext[out_ports[j]];
  }
  auto os = c_strides(oext.data(), out_rank);
  int64_t base = 0;
  std::vector<LDim> dims(src_rank);
  for (int i = 0; i < src_rank; i++) {
    dims[i].label = i; dims[i].ext = sext[i]; dims[i].sl = ss[i]; dims[i].in_l = true;
    if (fixed_idx && fixed_idx[i] >= 0) {
      TSB_CHECK(fixed_idx[i] < sext[i], TSB_ERR_INVALID, "index %d out of bounds for axis %d", fixed_idx[i], i);
      base += ss[i] * fixed_idx[i];
      dims[i].ext = 1;
    }
  }
  bool dups = false;
  for (int j = 0; j < out_rank; j++) {
    if (dims[out_ports[j]].in_o) dups = true;
    dims[out_ports[j]].so += os[j];
    dims[out_ports[j]].in_o = true;
  }
  const int64_t n_out = numel(oext.data(), out_rank);
  const uint8_t* sp = (const uint8_t*)src + base * dtype_size(ds);
  bool drops = false;
  for (int i = 0; i < src_rank; i++)
    if (!dims[i].in_o && dims[i].ext > 1) drops = true;
  // pure permutation of a dense binary tensor: the bit-permutation kernel
  if (!dups && !drops && !accumulate && ds == dd && !clamp && !fixed_idx && all_pow2(sext, src_rank)) {
    std::vector<int32_t> ident(src_rank), ex(src_rank);
    for (int i = 0; i < src_rank; i++) ident[i] = i;
    BitLabels sb = bit_labels(ident.data(), src_rank, sext);
    BitLabels db = bit_labels(out_ports, out_rank, sext);
    permute_bits(ctx, src, sb, dst, db, dtype_size(ds), stream);
    return;
  }
  if (dups && !accumulate) TSB_CUDA(cudaMemsetAsync(dst, 0, (size_t)n_out * dtype_size(dd), stream));
  GenDesc g = make_gen_desc(dims, ds, ds, dd, clamp, false, accumulate);
  launch_generic(ctx, g, sp, nullptr, dst, stream);
}

CaptureState& capture_state() {
  static thread_local CaptureState st;
  return st;
}

void ctx_unref(tsb_ctx* ctx) {
  if (!ctx || --ctx->refs > 0) return;
  if (ctx->d_flag) cudaFree(ctx->d_flag);
  delete ctx;
}

}  // namespace tsb

using namespace tsb;

extern "C" {

int tsb_version(void) { return TSB_VERSION; }
const char* tsb_last_error(void) { return tsb::last_error(); }

int tsb_ctx_create(int device, tsb_ctx** ctx) {
  TSB_API_BEGIN
  TSB_CHECK(ctx != nullptr, TSB_ERR_INVALID, "null output pointer");
  TSB_CUDA(cudaSetDevice(device));
  std::unique_ptr<tsb_ctx> c(new tsb_ctx());
  c->device = device;
  TSB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  TSB_CUDA(cudaMalloc((void**)&c->d_flag, 512));
  c->d_one = (uint8_t*)c->d_flag + 256;
  const uint8_t one = 1;
  TSB_CUDA(cudaMemcpy(c->d_one, &one, 1, cudaMemcpyHostToDevice));
  *ctx = c.release();
  TSB_API_END
}

int tsb_ctx_destroy(tsb_ctx* ctx) {
  TSB_API_BEGIN
  tsb::ctx_unref(ctx);   // freed once the plans created from it are gone too
  TSB_API_END
}

int tsb_ctx_launch_count(tsb_ctx* ctx, int64_t* count) {
  TSB_API_BEGIN
  check_ctx(ctx);
  *count = ctx->launches;
  TSB_API_END
}

int tsb_ctx_trace(tsb_ctx* ctx, int enabled) {
  TSB_API_BEGIN
  check_ctx(ctx);
  ctx->trace = enabled != 0;
  ctx->kcount.clear();
  TSB_API_END
}

int tsb_ctx_trace_json(tsb_ctx* ctx, char* buf, size_t cap, size_t* needed) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(needed != nullptr, TSB_ERR_INVALID, "null argument");
  std::string s = "{";
  for (auto& kv : ctx->kcount) {
    if (s.size() > 1) s += ",";
    s += "\"" + kv.first + "\":" + std::to_string(kv.second);
  }
  s += "}";
  *needed = s.size() + 1;
  if (buf && cap >= s.size() + 1) memcpy(buf, s.c_str(), s.size() + 1);
  TSB_API_END
}

int tsb_arena_reserve(tsb_ctx* ctx, size_t bytes) {
  TSB_API_BEGIN
  check_ctx(ctx);
  void* p = ctx->arena.alloc(bytes);
  ctx->arena.release(p);
  TSB_API_END
}
int tsb_arena_alloc(tsb_ctx* ctx, size_t bytes, void** ptr) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(ptr != nullptr, TSB_ERR_INVALID, "null output pointer");
  *ptr = ctx->arena.alloc(bytes);
  TSB_API_END
}
int tsb_arena_release(tsb_ctx* ctx, void* ptr) {
  TSB_API_BEGIN
  check_ctx(ctx);
  ctx->arena.release(ptr);
  TSB_API_END
}
int tsb_arena_alloc_async(tsb_ctx* ctx, size_t bytes, tsb_stream stream, void** ptr) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(ptr != nullptr, TSB_ERR_INVALID, "null output pointer");
  *ptr = ctx->arena.alloc(bytes, (cudaStream_t)stream);
  TSB_API_END
}
int tsb_arena_release_async(tsb_ctx* ctx, void* ptr, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  ctx->arena.release(ptr, (cudaStream_t)stream);
  TSB_API_END
}
int tsb_arena_trim(tsb_ctx* ctx) {
  TSB_API_BEGIN
  check_ctx(ctx);
  ctx->arena.trim();
  TSB_API_END
}
int tsb_arena_stats(tsb_ctx* ctx, tsb_arena_stats_t* stats) {
  TSB_API_BEGIN
  check_ctx(ctx);
  *stats = ctx->arena.st;
  TSB_API_END
}

int tsb_permute(tsb_ctx* ctx, const void* src, void* dst, int rank, const int32_t* extents, const int32_t* perm,
                int elem_size, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(rank >= 0 && rank <= TSB_MAX_RANK, TSB_ERR_INVALID, "rank out of range");
  TSB_CHECK(elem_size == 1 || elem_size == 4 || elem_size == 8, TSB_ERR_INVALID, "element size must be 1, 4 or 8");
  std::vector<char> seen(rank, 0);
  for (int i = 0; i < rank; i++) {
    TSB_CHECK(perm[i] >= 0 && perm[i] < rank && !seen[perm[i]], TSB_ERR_INVALID, "perm is not a permutation");
    seen[perm[i]] = 1;
  }
  const int dt = elem_size == 1 ? TSB_U8 : elem_size == 4 ? TSB_I32 : TSB_I64;
  rewire_impl(ctx, src, rank, extents, nullptr, dt, dst, rank, perm, dt, false, false, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_rewire(tsb_ctx* ctx, const void* src, int src_rank, const int32_t* src_extents, int src_dtype, void* dst,
               int out_rank, const int32_t* out_ports, int dst_dtype, int mode, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_mode(mode);
  rewire_impl(ctx, src, src_rank, src_extents, nullptr, src_dtype, dst, out_rank, out_ports, dst_dtype,
              mode == TSB_SAT, false, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_contract2(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires, int lhs_dtype,
                  const void* rhs, int rhs_rank, const int32_t* rhs_wires, int rhs_dtype, void* out, int out_rank,
                  const int32_t* out_wires, int out_dtype, int n_wires, const int32_t* wire_extents, int mode,
                  tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  contract2_impl(ctx, lhs, lhs_rank, lhs_wires, lhs_dtype, rhs, rhs_rank, rhs_wires, rhs_dtype, out, out_rank,
                 out_wires, out_dtype, n_wires, wire_extents, mode, false, false, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_slice(tsb_ctx* ctx, const void* src, int rank, const int32_t* extents, const int32_t* fixed_idx,
              int out_rank, const int32_t* out_axes, void* dst, int elem_size, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(elem_size == 1 || elem_size == 4 || elem_size == 8, TSB_ERR_INVALID, "element size must be 1, 4 or 8");
  const int dt = elem_size == 1 ? TSB_U8 : elem_size == 4 ? TSB_I32 : TSB_I64;
  int kept = 0;
  std::vector<char> seen(rank, 0);
  for (int i = 0; i < rank; i++) kept += fixed_idx[i] < 0;
  TSB_CHECK(kept == out_rank, TSB_ERR_INVALID, "slice: out_axes must list every kept axis exactly once");
  for (int j = 0; j < out_rank; j++) {
    TSB_CHECK(out_axes[j] >= 0 && out_axes[j] < rank && !seen[out_axes[j]], TSB_ERR_INVALID,
              "slice: out_axes must list every kept axis exactly once");
    seen[out_axes[j]] = 1;
  }
  rewire_impl(ctx, src, rank, extents, fixed_idx, dt, dst, out_rank, out_axes, dt, false, false, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_stack(tsb_ctx* ctx, const void* const* srcs, int n, int rank, const int32_t* extents, int axis, void* dst,
              int elem_size, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(n >= 1, TSB_ERR_INVALID, "Need at least one box to stack.");
  TSB_CHECK(axis >= 0 && axis <= rank, TSB_ERR_INVALID, "Invalid port index %d for stacking.", axis);
  TSB_CHECK(elem_size == 1 || elem_size == 4 || elem_size == 8, TSB_ERR_INVALID, "element size must be 1, 4 or 8");
  const int dt = elem_size == 1 ? TSB_U8 : elem_size == 4 ? TSB_I32 : TSB_I64;
  int64_t inner = 1, total = numel(extents, rank);
  for (int i = axis; i < rank; i++) inner *= extents[i];
  const int64_t outer = inner ? total / inner : 0;
  for (int k = 0; k < n; k++) {
    uint8_t* d = (uint8_t*)dst + (int64_t)k * inner * elem_size;
    if (outer == 1) {
      TSB_CUDA(cudaMemcpyAsync(d, srcs[k], (size_t)total * elem_size, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
      continue;
    }
    std::vector<LDim> dims(2);
    dims[0].label = 0; dims[0].ext = outer; dims[0].sl = inner; dims[0].so = inner * n; dims[0].in_l = dims[0].in_o = true;
    dims[1].label = 1; dims[1].ext = inner; dims[1].sl = 1; dims[1].so = 1; dims[1].in_l = dims[1].in_o = true;
    GenDesc g = make_gen_desc(dims, dt, dt, dt, false, false, false);
    launch_generic(ctx, g, srcs[k], nullptr, d, (cudaStream_t)stream);
  }
  TSB_API_END
}

int tsb_union(tsb_ctx* ctx, const void* a, const void* b, void* out, int64_t n, int dtype, int mode,
              tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype); check_mode(mode);
  TSB_CHECK(n >= 0, TSB_ERR_INVALID, "negative size");
  if (n == 0) return TSB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  (void)s;
  if (mode == TSB_SAT) {
    TSB_CHECK(dtype == TSB_U8, TSB_ERR_INVALID, "boolean union expects u8 tensors");
    const bool al = ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0);
    if (al) {
      TSB_LAUNCH(k_or_u8, dim3(grid_for(ctx, (n + 15) / 16, 256)), dim3(256), 0, s, (const uint8_t*)a,
                 (const uint8_t*)b, (uint8_t*)out, n);
    } else {
      // unaligned views: scalar path of the same kernel (no vector body)
      TSB_LAUNCH(k_add<uint8_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (const uint8_t*)a, (const uint8_t*)b,
                 (uint8_t*)out, n);
      TSB_LAUNCH(k_convert, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (const void*)out, (int)TSB_U8, out,
                 (int)TSB_U8, n, 1);
      ctx->note("k_convert");
    }
  } else if (dtype == TSB_U8) {
    TSB_LAUNCH(k_add<uint8_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (const uint8_t*)a, (const uint8_t*)b,
               (uint8_t*)out, n);
  } else if (dtype == TSB_I32) {
    TSB_LAUNCH(k_add<int32_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (const int32_t*)a, (const int32_t*)b,
               (int32_t*)out, n);
  } else {
    TSB_LAUNCH(k_add<int64_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (const int64_t*)a, (const int64_t*)b,
               (int64_t*)out, n);
  }
  ctx->note("k_union");
  TSB_CUDA(cudaGetLastError());
  TSB_API_END
}

static void fill_impl(tsb_ctx* ctx, void* dst, int64_t n, int64_t value, int dtype, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == TSB_U8) TSB_LAUNCH(k_fill<uint8_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (uint8_t*)dst, n, (uint8_t)value);
  else if (dtype == TSB_I32) TSB_LAUNCH(k_fill<int32_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (int32_t*)dst, n, (int32_t)value);
  else TSB_LAUNCH(k_fill<int64_t>, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, (int64_t*)dst, n, (int64_t)value);
  ctx->note("k_fill");
  TSB_CUDA(cudaGetLastError());
}

int tsb_fill(tsb_ctx* ctx, void* dst, int64_t n, int64_t value, int dtype, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype);
  fill_impl(ctx, dst, n, value, dtype, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_spider(tsb_ctx* ctx, void* dst, int dim, int num_ports, int dtype, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype);
  TSB_CHECK(dim >= 1 && num_ports >= 0, TSB_ERR_INVALID, "invalid spider arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t n = 1, diag = 0, st = 1;
  for (int i = 0; i < num_ports; i++) {
    diag += st;
    st *= dim;
    n *= dim;
    TSB_CHECK(n < (int64_t(1) << 40), TSB_ERR_INVALID, "spider too large");
  }
  if (num_ports == 0) {
    fill_impl(ctx, dst, 1, 1, dtype, s);
    return TSB_OK;
  }
  TSB_CUDA(cudaMemsetAsync(dst, 0, (size_t)n * dtype_size(dtype), s));
  const int grid = (dim + 127) / 128;
  if (dtype == TSB_U8) TSB_LAUNCH(k_diag_ones<uint8_t>, dim3(grid), dim3(128), 0, s, (uint8_t*)dst, dim, diag);
  else if (dtype == TSB_I32) TSB_LAUNCH(k_diag_ones<int32_t>, dim3(grid), dim3(128), 0, s, (int32_t*)dst, dim, diag);
  else TSB_LAUNCH(k_diag_ones<int64_t>, dim3(grid), dim3(128), 0, s, (int64_t*)dst, dim, diag);
  ctx->note("k_diag_ones");
  TSB_CUDA(cudaGetLastError());
  TSB_API_END
}

int tsb_convert(tsb_ctx* ctx, const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, int clamp,
                tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(src_dtype); check_dtype(dst_dtype);
  if (n <= 0) return TSB_OK;
  TSB_LAUNCH(k_convert, dim3(grid_for(ctx, n, 256)), dim3(256), 0, (cudaStream_t)stream, src, src_dtype, dst,
             dst_dtype, n, clamp);
  ctx->note("k_convert");
  TSB_CUDA(cudaGetLastError());
  TSB_API_END
}

int tsb_pack_bits(tsb_ctx* ctx, const void* src_u8, void* dst_bits, int64_t n, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(src_u8 && dst_bits && n >= 1, TSB_ERR_INVALID, "pack_bits: null buffer or empty tensor");
  TSB_CHECK(((uintptr_t)src_u8 & 15) == 0 && ((uintptr_t)dst_bits & 15) == 0, TSB_ERR_INVALID,
            "pack_bits: buffers must be 16-byte aligned");
  launch_pack_bits(ctx, src_u8, dst_bits, n, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_unpack_bits(tsb_ctx* ctx, const void* src_bits, void* dst_u8, int64_t n, int accumulate, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(src_bits && dst_u8 && n >= 1, TSB_ERR_INVALID, "unpack_bits: null buffer or empty tensor");
  TSB_CHECK(((uintptr_t)src_bits & 15) == 0 && ((uintptr_t)dst_u8 & 15) == 0, TSB_ERR_INVALID,
            "unpack_bits: buffers must be 16-byte aligned");
  launch_unpack_bits(ctx, src_bits, dst_u8, n, accumulate != 0, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_contract2_bits(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires, const void* rhs,
                       int rhs_rank, const int32_t* rhs_wires, void* out, int out_rank, const int32_t* out_wires,
                       int n_wires, const int32_t* wire_extents, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  TSB_CHECK(lhs && rhs && out && lhs_wires && rhs_wires && out_wires && wire_extents, TSB_ERR_INVALID, "null argument");
  auto chk = [&](const int32_t* ws, int r) {
    TSB_CHECK(r >= 0 && r <= TSB_MAX_RANK, TSB_ERR_INVALID, "rank out of range");
    for (int i = 0; i < r; i++) {
      TSB_CHECK(ws[i] >= 0 && ws[i] < n_wires, TSB_ERR_INVALID, "wire %d out of range", ws[i]);
      TSB_CHECK(wire_extents[ws[i]] == 2, TSB_ERR_UNSUPPORTED, "bit-packed contraction needs extent 2 on every wire");
    }
  };
  chk(lhs_wires, lhs_rank); chk(rhs_wires, rhs_rank); chk(out_wires, out_rank);
  BitLabels lb = bit_labels(lhs_wires, lhs_rank, wire_extents);
  BitLabels rb = bit_labels(rhs_wires, rhs_rank, wire_extents);
  BitLabels ob = bit_labels(out_wires, out_rank, wire_extents);
  MergeArgs m = make_merge_desc(lb, &rb, ob, false, 7);
  TSB_CHECK(m.ok && m.d.kbits <= 8, TSB_ERR_UNSUPPORTED,
            "bit-packed contraction needs one wire order for all three tensors, out_rank >= 7 and <= 8 summed wires");
  launch_merge_bits(ctx, m.d, lhs, 0, rhs, 0, out, (cudaStream_t)stream);
  TSB_API_END
}

int tsb_all_ones(tsb_ctx* ctx, const void* ptr, int64_t n, int dtype, int* result, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype);
  cudaStream_t s = (cudaStream_t)stream;
  int one = 1;
  TSB_CUDA(cudaMemcpyAsync(ctx->d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, s));
  if (n > 0) {
    TSB_LAUNCH(k_all_ones, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, ptr, dtype, n, ctx->d_flag);
    ctx->note("k_all_ones");
  }
  TSB_CUDA(cudaMemcpyAsync(result, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  TSB_CUDA(cudaStreamSynchronize(s));
  TSB_API_END
}

int tsb_all_binary(tsb_ctx* ctx, const void* ptr, int64_t n, int dtype, int* result, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype);
  cudaStream_t s = (cudaStream_t)stream;
  int one = 1;
  TSB_CUDA(cudaMemcpyAsync(ctx->d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, s));
  if (n > 0) {
    TSB_LAUNCH(k_all_binary, dim3(grid_for(ctx, n, 256)), dim3(256), 0, s, ptr, dtype, n, ctx->d_flag);
    ctx->note("k_all_binary");
  }
  TSB_CUDA(cudaMemcpyAsync(result, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  TSB_CUDA(cudaStreamSynchronize(s));
  TSB_API_END
}

int tsb_scalar_to_host(tsb_ctx* ctx, const void* ptr, int dtype, int64_t* out, tsb_stream stream) {
  TSB_API_BEGIN
  check_ctx(ctx);
  check_dtype(dtype);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char buf[8] = {0};
  TSB_CUDA(cudaMemcpyAsync(buf, ptr, dtype_size(dtype), cudaMemcpyDeviceToHost, s));
  TSB_CUDA(cudaStreamSynchronize(s));
  if (dtype == TSB_U8) *out = buf[0];
  else if (dtype == TSB_I32) { int32_t v; memcpy(&v, buf, 4); *out = v; }
  else { int64_t v; memcpy(&v, buf, 8); *out = v; }
  TSB_API_END
}

}  // extern "C"
