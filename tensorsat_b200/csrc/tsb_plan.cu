// tsb_plan.cu — whole-tree plans: the device-resident stand-in for
// ``tree.contract(diagram.boxes, prefer_einsum=True)`` as driven by CotengraContraction.contract
// (tensorsat/contractions/cotengra.py:116-128) and for SimpleContraction.contract
// (tensorsat/contractions/simple.py:251-292).
//
// Design:
//  * one global label order for every tensor of the plan — sliced wires on top, then wires by the
//    step at which they are summed (soonest = most significant), output wires at the bottom —
//    so every step is a monotone merge and no intermediate is ever transposed;
//  * lazy slicing: tensors that stay small with the sliced wires kept are computed ONCE for all
//    slices ("shared" part); only the large top of the tree runs per slice, reading shared
//    operands through a pointer offset;
//  * static memory plan: every intermediate has a known lifetime, so offsets into one arena
//    block are assigned at plan time (peak = tsb_plan_info().peak_bytes);
//  * per-tensor dtype in counting mode from a magnitude bound (u8 -> i32 -> i64).
#include <cmath>
#include <functional>
#include <sstream>

#include "lower.h"

namespace tsb {
void contract2_impl(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires, int dl, const void* rhs,
                    int rhs_rank, const int32_t* rhs_wires, int dr, void* out, int out_rank,
                    const int32_t* out_wires, int dout, int n_wires, const int32_t* wext, int mode, bool accumulate,
                    bool force_generic, cudaStream_t stream);
void rewire_impl(tsb_ctx* ctx, const void* src, int src_rank, const int32_t* sext, const int32_t* fixed_idx, int ds,
                 void* dst, int out_rank, const int32_t* out_ports, int dd, bool clamp, bool accumulate,
                 cudaStream_t stream);

struct PNode {
  std::vector<int32_t> slab;     // sliced wires present below this node (order of desc.sliced_wires)
  std::vector<int32_t> reg;      // regular wires, global order (most significant first)
  bool shared = false;           // computed once, sliced wires kept as leading axes
  int dtype = TSB_U8;
  double bound_log2 = 0;         // log2 of the largest possible entry (counting mode)
  int64_t reg_numel = 1;         // product of regular extents
  int64_t s_numel = 1;           // product of present sliced extents
  int64_t offset = -1;           // byte offset in the arena block (-1: not allocated / external)
  size_t bytes = 0;
  bool is_out = false;           // the root, written straight into the caller's buffer
  std::vector<int32_t> stored() const {  // axis order of the stored buffer
    std::vector<int32_t> v;
    if (shared) v = slab;
    v.insert(v.end(), reg.begin(), reg.end());
    return v;
  }
};

enum OpKind { OP_LEAF = 0, OP_STEP = 1, OP_FINAL_PERMUTE = 2 };

struct POp {
  int kind = OP_STEP;
  bool per_slice = false;
  int lhs = -1, rhs = -1, out = -1;  // node ids
  bool use_merge = false;
  MergeDesc md{};
  // generic description (wires + extents) used when the merge kernel does not apply
  std::vector<int32_t> lw, rw, ow;
  // leaf preparation
  int leaf = -1;
  std::vector<int32_t> leaf_ports;  // dst axis j <- leaf axis leaf_ports[j]
  double alg_bytes = 0, alg_ops = 0;
  int b = 0, m = 0, n = 0, k = 0;  // label class sizes in bits (for reports)
  float last_ms = 0;
};

}  // namespace tsb

using namespace tsb;

struct tsb_plan {
  tsb_ctx* ctx = nullptr;
  int mode = TSB_SAT;
  uint32_t flags = 0;
  int n_leaves = 0, n_wires = 0;
  std::vector<std::vector<int32_t>> leaf_wires;
  std::vector<int32_t> out_wires, wext, sliced;
  std::vector<PNode> nodes;
  std::vector<POp> ops;
  tsb_plan_info_t info{};
  int root = -1;
  bool stack_out = false;       // some output wire is sliced: per-slice results land in `rbuf`
  int rnode = -1;               // pseudo node for rbuf
  std::vector<int32_t> rbuf_axes;
  bool timing = false;
  std::vector<cudaEvent_t> ev;
};

namespace {

struct FreeList {  // first-fit offset allocator over one block
  std::map<int64_t, int64_t> free_;  // offset -> size
  int64_t top = 0, peak = 0;
  static int64_t al(int64_t n) { return (n + 255) / 256 * 256; }
  int64_t alloc(int64_t bytes) {
    const int64_t need = al(std::max<int64_t>(bytes, 1));
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= need) {
        const int64_t off = it->first, rest = it->second - need;
        free_.erase(it);
        if (rest > 0) free_[off + need] = rest;
        return off;
      }
    }
    // grow: extend a trailing free block if there is one
    if (!free_.empty()) {
      auto last = std::prev(free_.end());
      if (last->first + last->second == top) {
        const int64_t off = last->first;
        free_.erase(last);
        top = off + need;
        peak = std::max(peak, top);
        return off;
      }
    }
    const int64_t off = top;
    top += need;
    peak = std::max(peak, top);
    return off;
  }
  void release(int64_t off, int64_t bytes) {
    int64_t size = al(std::max<int64_t>(bytes, 1));
    auto nx = free_.lower_bound(off);
    if (nx != free_.end() && off + size == nx->first) {
      size += nx->second;
      nx = free_.erase(nx);
    }
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) {
        pv->second += size;
        return;
      }
    }
    free_[off] = size;
  }
};

int dtype_for_bound(double b) { return b <= 7.0 ? TSB_U8 : b <= 30.0 ? TSB_I32 : TSB_I64; }

int64_t prod_ext(const std::vector<int32_t>& ws, const std::vector<int32_t>& wext) {
  int64_t n = 1;
  for (auto w : ws) n *= wext[w];
  return n;
}

void build_plan(tsb_plan& P, const tsb_plan_desc& D) {
  tsb_ctx* ctx = P.ctx;
  (void)ctx;
  TSB_CHECK(D.n_leaves >= 1, TSB_ERR_INVALID, "Cannot contract an empty network.");
  TSB_CHECK(D.n_wires >= 0 && D.n_out >= 0 && D.n_sliced >= 0, TSB_ERR_INVALID, "negative count in plan description");
  TSB_CHECK(D.mode == TSB_SAT || D.mode == TSB_COUNT, TSB_ERR_INVALID, "invalid mode %d", D.mode);
  P.mode = D.mode;
  P.flags = D.flags;
  P.n_leaves = D.n_leaves;
  P.n_wires = D.n_wires;
  P.wext.assign(D.wire_extents, D.wire_extents + D.n_wires);
  for (auto e : P.wext) TSB_CHECK(e >= 1, TSB_ERR_INVALID, "wire extent must be positive");
  const int L = D.n_leaves;
  auto chkw = [&](int32_t w) { TSB_CHECK(w >= 0 && w < D.n_wires, TSB_ERR_INVALID, "wire %d out of range", w); };
  {
    const int32_t* p = D.leaf_wires;
    for (int i = 0; i < L; i++) {
      TSB_CHECK(D.leaf_ranks[i] >= 0 && D.leaf_ranks[i] <= TSB_MAX_RANK, TSB_ERR_INVALID, "leaf rank out of range");
      P.leaf_wires.emplace_back(p, p + D.leaf_ranks[i]);
      for (auto w : P.leaf_wires.back()) chkw(w);
      p += D.leaf_ranks[i];
    }
  }
  P.out_wires.assign(D.out_wires, D.out_wires + D.n_out);
  P.sliced.assign(D.sliced_wires, D.sliced_wires + D.n_sliced);
  for (auto w : P.out_wires) chkw(w);
  for (auto w : P.sliced) chkw(w);
  {
    std::set<int32_t> so(P.out_wires.begin(), P.out_wires.end()), ss(P.sliced.begin(), P.sliced.end());
    TSB_CHECK(so.size() == P.out_wires.size(), TSB_ERR_INVALID, "plan output wires must be distinct");
    TSB_CHECK(ss.size() == P.sliced.size(), TSB_ERR_INVALID, "Sliced wires must be distinct.");
  }
  std::set<int32_t> S(P.sliced.begin(), P.sliced.end()), OUT(P.out_wires.begin(), P.out_wires.end());
  std::map<int32_t, int> s_rank;
  for (int i = 0; i < (int)P.sliced.size(); i++) s_rank[P.sliced[i]] = i;

  // ---- label bookkeeping with the "last two holders" rule (hyper-edges) -------------------
  std::vector<int> count(D.n_wires, 0);
  std::vector<std::set<int32_t>> regset(2 * L - 1), sset(2 * L - 1);
  for (int i = 0; i < L; i++)
    for (auto w : P.leaf_wires[i]) (S.count(w) ? sset[i] : regset[i]).insert(w);
  for (int i = 0; i < L; i++)
    for (auto w : regset[i]) count[w]++;
  std::vector<char> seen_any(D.n_wires, 0);
  for (int i = 0; i < L; i++)
    for (auto w : P.leaf_wires[i]) seen_any[w] = 1;
  for (auto w : P.out_wires) {
    TSB_CHECK(seen_any[w], TSB_ERR_INVALID, "output wire %d is not attached to any tensor", w);
    if (!S.count(w)) count[w]++;
  }
  // lone wires: on one leaf only and not an output -> summed during leaf preparation
  std::vector<std::set<int32_t>> lone(L);
  std::vector<double> bound(2 * L - 1, 0.0);
  auto freew = [&](int32_t w) { return D.wire_is_free ? (D.wire_is_free[w] != 0) : true; };
  for (int i = 0; i < L; i++) {
    std::vector<int32_t> drop;
    for (auto w : regset[i])
      if (count[w] == 1) drop.push_back(w);
    for (auto w : drop) {
      regset[i].erase(w);
      lone[i].insert(w);
      count[w] = 0;
      if (freew(w)) bound[i] += std::log2((double)P.wext[w]);
    }
  }
  const int n_steps = L - 1;
  std::vector<int> death(D.n_wires, INT32_MAX);
  std::vector<std::pair<int, int>> path(n_steps);
  std::vector<char> alive(2 * L - 1, 0);
  for (int i = 0; i < L; i++) alive[i] = 1;
  for (int s = 0; s < n_steps; s++) {
    const int a = D.ssa_path[2 * s], b = D.ssa_path[2 * s + 1];
    TSB_CHECK(a >= 0 && a < L + s && b >= 0 && b < L + s && a != b && alive[a] && alive[b], TSB_ERR_INVALID,
              "Invalid contraction pair (%d, %d) at step %d.", a, b, s);
    path[s] = {a, b};
    alive[a] = alive[b] = 0;
    alive[L + s] = 1;
    std::set<int32_t> u = regset[a];
    u.insert(regset[b].begin(), regset[b].end());
    double bd = bound[a] + bound[b];
    for (auto w : u) {
      const int c = count[w] - ((regset[a].count(w) && regset[b].count(w)) ? 1 : 0);
      if (c >= 2) {
        regset[L + s].insert(w);
        count[w] = c;
      } else {
        count[w] = 0;
        death[w] = s;
        if (freew(w)) bd += std::log2((double)P.wext[w]);
      }
    }
    bound[L + s] = bd;
    sset[L + s] = sset[a];
    sset[L + s].insert(sset[b].begin(), sset[b].end());
  }
  P.root = 2 * L - 2;
  if (n_steps > 0) {
    // the root must hold exactly the non-sliced output wires
    std::set<int32_t> want;
    for (auto w : P.out_wires) if (!S.count(w)) want.insert(w);
    TSB_CHECK(regset[P.root] == want, TSB_ERR_INVALID, "internal: root labels do not match the output wires");
  }

  // ---- global order: rank[w] smaller = more significant ------------------------------------
  std::vector<int64_t> rank(D.n_wires, 0);
  {
    std::map<int32_t, int> out_pos;
    for (int i = 0; i < (int)P.out_wires.size(); i++) out_pos[P.out_wires[i]] = i;
    for (int w = 0; w < D.n_wires; w++) {
      if (OUT.count(w) && !S.count(w)) rank[w] = (int64_t(1) << 40) + out_pos[w];
      else rank[w] = (int64_t)death[w] * 4096 * 1024 + w;  // by death step, then wire id
      if (death[w] == INT32_MAX && !OUT.count(w)) rank[w] = (int64_t(1) << 39) + w;  // lone / unused
    }
  }
  auto order_reg = [&](const std::set<int32_t>& s) {
    std::vector<int32_t> v(s.begin(), s.end());
    std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) { return rank[a] < rank[b]; });
    return v;
  };
  auto order_s = [&](const std::set<int32_t>& s) {
    std::vector<int32_t> v(s.begin(), s.end());
    std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) { return s_rank[a] < s_rank[b]; });
    return v;
  };

  // ---- nodes ---------------------------------------------------------------------------------
  const int batch_log2 = D.batch_log2 >= 0 ? D.batch_log2 : 18;
  P.nodes.resize(2 * L - 1);
  const bool sliced_any = !P.sliced.empty();
  for (int i = 0; i < 2 * L - 1; i++) {
    PNode& nd = P.nodes[i];
    nd.reg = order_reg(regset[i]);
    nd.slab = order_s(sset[i]);
    nd.reg_numel = prod_ext(nd.reg, P.wext);
    nd.s_numel = prod_ext(nd.slab, P.wext);
    nd.bound_log2 = bound[i];
    nd.dtype = P.mode == TSB_SAT ? TSB_U8 : dtype_for_bound(bound[i]);
    if (!sliced_any || i < L) {
      nd.shared = true;  // nothing to slice / leaves are handed over unsliced
    } else {
      const auto& pr = path[i - L];
      const double lg = std::log2((double)nd.reg_numel) + std::log2((double)nd.s_numel);
      nd.shared = P.nodes[pr.first].shared && P.nodes[pr.second].shared && lg <= batch_log2 && i != P.root;
    }
  }
  if (P.mode == TSB_COUNT) {
    P.nodes[P.root].dtype = TSB_I64;
    double tot = bound[P.root];
    for (auto w : P.sliced) if (freew(w)) tot += std::log2((double)P.wext[w]);
    P.info.may_overflow = tot > 62.0;
    for (int i = 0; i < 2 * L - 1; i++) if (bound[i] > 62.0) P.info.may_overflow = 1;
  }
  // sliced output wires: per-slice results are stacked into an intermediate buffer
  for (auto w : P.out_wires) if (S.count(w)) P.stack_out = true;

  // ---- ops -----------------------------------------------------------------------------------
  auto esz = [&](int node) { return (int64_t)dtype_size(P.nodes[node].dtype); };
  for (int i = 0; i < L; i++) {
    POp op;
    op.kind = OP_LEAF;
    op.leaf = i;
    op.out = i;
    op.per_slice = false;
    const auto& lw = P.leaf_wires[i];
    for (auto w : P.nodes[i].stored()) {
      int pos = -1;
      for (int q = 0; q < (int)lw.size(); q++) if (lw[q] == w) { pos = q; break; }
      op.leaf_ports.push_back(pos);
    }
    op.alg_bytes = 0;  // leaf staging is not algorithmic work
    P.ops.push_back(op);
  }
  if (L == 1) {
    // single tensor: the "contraction" is the rewire of the leaf onto the output wires
    P.nodes[0].is_out = !P.stack_out && !sliced_any;
  }
  for (int s = 0; s < n_steps; s++) {
    POp op;
    op.kind = OP_STEP;
    op.lhs = path[s].first;
    op.rhs = path[s].second;
    op.out = L + s;
    PNode& o = P.nodes[op.out];
    op.per_slice = !o.shared;
    // put the larger operand on the lhs for reporting symmetry (no semantic effect)
    const PNode& a = P.nodes[op.lhs];
    const PNode& b = P.nodes[op.rhs];
    // wires seen by the kernel: shared ops see sliced wires as ordinary leading axes
    op.lw = o.shared ? a.stored() : a.reg;
    op.rw = o.shared ? b.stored() : b.reg;
    op.ow = o.shared ? o.stored() : o.reg;
    std::set<int32_t> sl(op.lw.begin(), op.lw.end()), sr(op.rw.begin(), op.rw.end()), so(op.ow.begin(), op.ow.end());
    double lb = 0, lm = 0, ln = 0, lk = 0;
    for (auto w : sl) {
      const double e = std::log2((double)P.wext[w]);
      if (sr.count(w)) (so.count(w) ? lb : lk) += e; else (so.count(w) ? lm : lk) += e;
    }
    for (auto w : sr) if (!sl.count(w)) { const double e = std::log2((double)P.wext[w]); (so.count(w) ? ln : lk) += e; }
    op.b = (int)lb; op.m = (int)lm; op.n = (int)ln; op.k = (int)lk;
    const double nl = (double)prod_ext(op.lw, P.wext), nr = (double)prod_ext(op.rw, P.wext),
                 no = (double)prod_ext(op.ow, P.wext);
    op.alg_bytes = nl * esz(op.lhs) + nr * esz(op.rhs) + no * esz(op.out);
    op.alg_ops = 2.0 * std::exp2(lb + lm + ln + lk);
    P.ops.push_back(op);
  }
  if (n_steps > 0 && !P.stack_out) P.nodes[P.root].is_out = true;

  // fast-kernel eligibility of each step
  const bool force_generic = (P.flags & TSB_PLAN_FORCE_GENERIC) != 0;
  for (auto& op : P.ops) {
    if (op.kind != OP_STEP || force_generic) continue;
    bool pow2 = true;
    for (auto w : op.lw) pow2 = pow2 && is_pow2(P.wext[w]);
    for (auto w : op.rw) pow2 = pow2 && is_pow2(P.wext[w]);
    if (!pow2) continue;
    BitLabels lb = bit_labels(op.lw.data(), (int)op.lw.size(), P.wext.data());
    BitLabels rb = bit_labels(op.rw.data(), (int)op.rw.size(), P.wext.data());
    BitLabels ob = bit_labels(op.ow.data(), (int)op.ow.size(), P.wext.data());
    MergeArgs m = make_merge_desc(lb, &rb, ob, false);
    if (m.ok) {
      op.use_merge = true;
      op.md = m.d;
    }
  }

  // ---- memory plan -----------------------------------------------------------------------------
  // timeline: shared ops in order, then the per-slice ops in order (repeated for every slice)
  std::vector<int> order;
  for (int i = 0; i < (int)P.ops.size(); i++) if (!P.ops[i].per_slice) order.push_back(i);
  for (int i = 0; i < (int)P.ops.size(); i++) if (P.ops[i].per_slice) order.push_back(i);
  std::vector<int> last_use(P.nodes.size(), -1);
  for (int t = 0; t < (int)order.size(); t++) {
    const POp& op = P.ops[order[t]];
    if (op.lhs >= 0) last_use[op.lhs] = t;
    if (op.rhs >= 0) last_use[op.rhs] = t;
  }
  // shared tensors consumed by per-slice ops must survive every slice iteration
  const int T_END = (int)order.size();
  for (auto& op : P.ops)
    if (op.per_slice) {
      if (op.lhs >= 0 && P.nodes[op.lhs].shared) last_use[op.lhs] = T_END;
      if (op.rhs >= 0 && P.nodes[op.rhs].shared) last_use[op.rhs] = T_END;
    }
  FreeList fl;
  for (int i = 0; i < (int)P.nodes.size(); i++) {
    PNode& nd = P.nodes[i];
    nd.bytes = (size_t)(nd.reg_numel * (nd.shared ? nd.s_numel : 1) * dtype_size(nd.dtype));
  }
  if (P.stack_out || (sliced_any && L == 1)) {
    // result buffer with the sliced output wires leading; combined across slices, permuted at the end
    PNode r;
    for (auto w : P.sliced) if (OUT.count(w)) r.slab.push_back(w);
    for (auto w : P.out_wires) if (!S.count(w)) r.reg.push_back(w);
    r.shared = true;
    r.dtype = P.mode == TSB_SAT ? TSB_U8 : TSB_I64;
    r.reg_numel = prod_ext(r.reg, P.wext);
    r.s_numel = prod_ext(r.slab, P.wext);
    r.bytes = (size_t)(r.reg_numel * r.s_numel * dtype_size(r.dtype));
    r.offset = fl.alloc((int64_t)r.bytes);
    P.rnode = (int)P.nodes.size();
    P.nodes.push_back(r);
    P.stack_out = true;
  }
  for (int t = 0; t < (int)order.size(); t++) {
    POp& op = P.ops[order[t]];
    PNode& o = P.nodes[op.out];
    if (!o.is_out) o.offset = fl.alloc((int64_t)o.bytes);
    for (int opnd : {op.lhs, op.rhs})
      if (opnd >= 0 && last_use[opnd] == t && P.nodes[opnd].offset >= 0) fl.release(P.nodes[opnd].offset, (int64_t)P.nodes[opnd].bytes);
  }
  P.info.peak_bytes = fl.peak;

  // ---- info ------------------------------------------------------------------------------------
  P.info.n_slices = prod_ext(P.sliced, P.wext);
  P.info.out_numel = prod_ext(P.out_wires, P.wext);
  P.info.out_dtype = P.mode == TSB_SAT ? TSB_U8 : TSB_I64;
  P.info.n_steps = n_steps;
  int mx = 0;
  for (auto& op : P.ops) {
    if (op.per_slice) {
      P.info.n_kernels_per_slice++;
      P.info.alg_bytes_per_slice += op.alg_bytes;
      P.info.alg_ops_per_slice += op.alg_ops;
    } else {
      P.info.n_kernels_shared++;
      P.info.alg_bytes_shared += op.alg_bytes;
      P.info.alg_ops_shared += op.alg_ops;
    }
    if (op.kind == OP_STEP) mx = std::max(mx, ilog2(std::max<int64_t>(1, prod_ext(op.ow, P.wext))));
  }
  P.info.log2_max_tensor = mx;
}

// digits of a slice id (first sliced wire most significant)
std::vector<int64_t> slice_digits(const tsb_plan& P, int64_t sid) {
  std::vector<int64_t> d(P.sliced.size());
  for (int i = (int)P.sliced.size() - 1; i >= 0; i--) {
    d[i] = sid % P.wext[P.sliced[i]];
    sid /= P.wext[P.sliced[i]];
  }
  return d;
}

// element offset of the slice's sub-block inside a shared node (sliced wires are its leading axes)
int64_t shared_offset(const tsb_plan& P, const PNode& nd, const std::vector<int64_t>& digits) {
  int64_t off = 0;
  for (auto w : nd.slab) {
    int idx = 0;
    for (int i = 0; i < (int)P.sliced.size(); i++) if (P.sliced[i] == w) idx = i;
    off = off * P.wext[w] + digits[idx];
  }
  return off * nd.reg_numel;
}

void run_step(tsb_plan& P, POp& op, uint8_t* base, void* out_ptr, const std::vector<int64_t>* digits, bool accumulate,
              cudaStream_t stream) {
  tsb_ctx* ctx = P.ctx;
  const PNode& a = P.nodes[op.lhs];
  const PNode& b = P.nodes[op.rhs];
  const PNode& o = P.nodes[op.out];
  const uint8_t* pl = base + a.offset;
  const uint8_t* pr = base + b.offset;
  if (op.per_slice) {
    if (a.shared) pl += shared_offset(P, a, *digits) * dtype_size(a.dtype);
    if (b.shared) pr += shared_offset(P, b, *digits) * dtype_size(b.dtype);
  }
  void* po = o.is_out ? out_ptr : (void*)(base + o.offset);
  bool acc = accumulate && o.is_out;
  if (P.stack_out && op.out == P.root) {
    // root result goes into the slice's block of the stacked result buffer, accumulating
    const PNode& r = P.nodes[P.rnode];
    po = base + r.offset + shared_offset(P, r, *digits) * dtype_size(r.dtype);
    acc = true;
  }
  const bool clamp = P.mode == TSB_SAT;
  int dout = o.dtype;
  if (P.stack_out && op.out == P.root) dout = P.nodes[P.rnode].dtype;
  if (op.use_merge && (((uintptr_t)pl | (uintptr_t)pr | (uintptr_t)po) & 15) == 0) {
    MergeDesc d = op.md;
    d.accumulate = acc ? 1 : 0;
    launch_merge(ctx, d, pl, a.dtype, pr, b.dtype, po, dout, clamp, stream);
  } else {
    contract2_impl(ctx, pl, (int)op.lw.size(), op.lw.data(), a.dtype, pr, (int)op.rw.size(), op.rw.data(), b.dtype, po,
                   (int)op.ow.size(), op.ow.data(), dout, P.n_wires, P.wext.data(), P.mode, acc,
                   (P.flags & TSB_PLAN_FORCE_GENERIC) != 0, stream);
  }
}

void run_leaf(tsb_plan& P, POp& op, uint8_t* base, const void* leaf_ptr, void* out_ptr, cudaStream_t stream) {
  // leaf -> plan order: an einsum against the scalar 1 (repeated wires = diagonal, lone wires = sum / or)
  const PNode& o = P.nodes[op.out];
  const auto& lw = P.leaf_wires[op.leaf];
  const std::vector<int32_t> ow = o.stored();
  void* po = o.is_out ? out_ptr : (void*)(base + o.offset);
  contract2_impl(P.ctx, leaf_ptr, (int)lw.size(), lw.data(), TSB_U8, P.ctx->d_one, 0, nullptr, TSB_U8, po,
                 (int)ow.size(), ow.data(), o.dtype, P.n_wires, P.wext.data(), P.mode, false, true, stream);
}

}  // namespace

extern "C" {

int tsb_plan_create(tsb_ctx* ctx, const tsb_plan_desc* desc, tsb_plan** plan) {
  try {
    TSB_CHECK(ctx && desc && plan, TSB_ERR_INVALID, "null argument");
    std::unique_ptr<tsb_plan> P(new tsb_plan());
    P->ctx = ctx;
    build_plan(*P, *desc);
    *plan = P.release();
    return TSB_OK;
  } catch (const tsb::Error& e) {
    return e.code;
  } catch (const std::exception& e) {
    tsb::set_error("internal error: %s", e.what());
    return TSB_ERR_INVALID;
  }
}

int tsb_plan_info(const tsb_plan* plan, tsb_plan_info_t* info) {
  if (!plan || !info) { tsb::set_error("null argument"); return TSB_ERR_INVALID; }
  *info = plan->info;
  return TSB_OK;
}

int tsb_plan_describe(const tsb_plan* plan, char* buf, size_t cap, size_t* needed) {
  if (!plan || !needed) { tsb::set_error("null argument"); return TSB_ERR_INVALID; }
  std::ostringstream os;
  os << "{\"peak_bytes\":" << plan->info.peak_bytes << ",\"n_slices\":" << plan->info.n_slices << ",\"ops\":[";
  bool first = true;
  for (auto& op : plan->ops) {
    if (!first) os << ",";
    first = false;
    const PNode& o = plan->nodes[op.out];
    os << "{\"kind\":\"" << (op.kind == OP_LEAF ? "leaf" : op.kind == OP_STEP ? (op.use_merge ? "merge" : "generic") : "permute")
       << "\",\"per_slice\":" << (op.per_slice ? 1 : 0) << ",\"out_bits\":" << ilog2(std::max<int64_t>(1, o.reg_numel))
       << ",\"dtype\":" << o.dtype << ",\"b\":" << op.b << ",\"m\":" << op.m << ",\"n\":" << op.n << ",\"k\":" << op.k
       << ",\"bytes\":" << op.alg_bytes << ",\"ops\":" << op.alg_ops << ",\"ms\":" << op.last_ms << "}";
  }
  os << "]}";
  const std::string s = os.str();
  *needed = s.size() + 1;
  if (buf && cap >= s.size() + 1) memcpy(buf, s.c_str(), s.size() + 1);
  return TSB_OK;
}

int tsb_plan_set_timing(tsb_plan* plan, int enabled) {
  if (!plan) { tsb::set_error("null argument"); return TSB_ERR_INVALID; }
  plan->timing = enabled != 0;
  return TSB_OK;
}

int tsb_plan_timing_json(tsb_plan* plan, char* buf, size_t cap, size_t* needed) {
  return tsb_plan_describe(plan, buf, cap, needed);
}

int tsb_plan_execute(tsb_plan* plan, const void* const* leaf_ptrs, int64_t slice_begin, int64_t slice_end, void* out,
                     tsb_stream stream_) {
  try {
    TSB_CHECK(plan && leaf_ptrs && out, TSB_ERR_INVALID, "null argument");
    tsb_plan& P = *plan;
    tsb_ctx* ctx = P.ctx;
    cudaStream_t stream = (cudaStream_t)stream_;
    TSB_CHECK(0 <= slice_begin && slice_begin <= slice_end && slice_end <= P.info.n_slices, TSB_ERR_INVALID,
              "slice range [%lld, %lld) out of [0, %lld]", (long long)slice_begin, (long long)slice_end,
              (long long)P.info.n_slices);
    const int es = dtype_size(P.info.out_dtype);
    if (slice_begin == slice_end) {  // nothing assigned to this rank: neutral element
      TSB_CUDA(cudaMemsetAsync(out, 0, (size_t)P.info.out_numel * es, stream));
      return TSB_OK;
    }
    Temp block(ctx, (size_t)std::max<int64_t>(P.info.peak_bytes, 256));
    uint8_t* base = (uint8_t*)block.p;
    const bool timing = P.timing;
    if (timing && P.ev.size() < 2 * P.ops.size() + 2) {
      size_t old = P.ev.size();
      P.ev.resize(2 * P.ops.size() + 2);
      for (size_t i = old; i < P.ev.size(); i++) TSB_CUDA(cudaEventCreate(&P.ev[i]));
    }
    for (auto& op : P.ops) op.last_ms = 0;
    auto timed = [&](size_t i, const std::function<void()>& f) {
      if (timing) TSB_CUDA(cudaEventRecord(P.ev[2 * i], stream));
      f();
      if (timing) {
        TSB_CUDA(cudaEventRecord(P.ev[2 * i + 1], stream));
        TSB_CUDA(cudaEventSynchronize(P.ev[2 * i + 1]));
        float ms = 0;
        TSB_CUDA(cudaEventElapsedTime(&ms, P.ev[2 * i], P.ev[2 * i + 1]));
        P.ops[i].last_ms += ms;
      }
    };
    if (P.stack_out) {
      const PNode& r = P.nodes[P.rnode];
      TSB_CUDA(cudaMemsetAsync(base + r.offset, 0, r.bytes, stream));
    }
    // shared part: once
    for (size_t i = 0; i < P.ops.size(); i++) {
      POp& op = P.ops[i];
      if (op.per_slice) continue;
      if (op.kind == OP_LEAF) {
        if (P.n_leaves == 1 && !P.sliced.empty()) continue;  // handled per slice below
        timed(i, [&] { run_leaf(P, op, base, leaf_ptrs[op.leaf], out, stream); });
      } else {
        timed(i, [&] { run_step(P, op, base, out, nullptr, false, stream); });
      }
    }
    // per-slice part
    if (!P.sliced.empty()) {
      for (int64_t sid = slice_begin; sid < slice_end; sid++) {
        const std::vector<int64_t> digits = slice_digits(P, sid);
        if (P.n_leaves == 1) {
          // single tensor network with sliced wires: index the leaf, accumulate into the stacked buffer
          const auto& lw = P.leaf_wires[0];
          std::vector<int32_t> ext(lw.size()), fixed(lw.size(), -1), ports;
          for (size_t q = 0; q < lw.size(); q++) {
            ext[q] = P.wext[lw[q]];
            for (size_t z = 0; z < P.sliced.size(); z++)
              if (P.sliced[z] == lw[q]) fixed[q] = (int32_t)digits[z];
          }
          const PNode& r = P.nodes[P.rnode];
          for (auto w : r.reg)
            for (size_t q = 0; q < lw.size(); q++) if (lw[q] == w) { ports.push_back((int32_t)q); break; }
          void* po = base + r.offset + shared_offset(P, r, digits) * dtype_size(r.dtype);
          rewire_impl(ctx, leaf_ptrs[0], (int)lw.size(), ext.data(), fixed.data(), TSB_U8, po, (int)ports.size(),
                      ports.data(), r.dtype, P.mode == TSB_SAT, true, stream);
          continue;
        }
        for (size_t i = 0; i < P.ops.size(); i++) {
          POp& op = P.ops[i];
          if (!op.per_slice) continue;
          timed(i, [&] { run_step(P, op, base, out, &digits, sid != slice_begin, stream); });
        }
      }
    }
    if (P.stack_out) {
      // [sliced output wires][other output wires] -> requested output order
      const PNode& r = P.nodes[P.rnode];
      std::vector<int32_t> axes = r.stored(), ext(axes.size()), ports;
      for (size_t q = 0; q < axes.size(); q++) ext[q] = P.wext[axes[q]];
      for (auto w : P.out_wires)
        for (size_t q = 0; q < axes.size(); q++) if (axes[q] == w) { ports.push_back((int32_t)q); break; }
      rewire_impl(ctx, base + r.offset, (int)axes.size(), ext.data(), nullptr, r.dtype, out, (int)ports.size(),
                  ports.data(), r.dtype, false, false, stream);
    }
    return TSB_OK;
  } catch (const tsb::Error& e) {
    return e.code;
  } catch (const std::exception& e) {
    tsb::set_error("internal error: %s", e.what());
    return TSB_ERR_INVALID;
  }
}

int tsb_plan_destroy(tsb_plan* plan) {
  if (plan) {
    for (auto e : plan->ev) cudaEventDestroy(e);
    delete plan;
  }
  return TSB_OK;
}

}  // extern "C"
