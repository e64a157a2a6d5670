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
#include <mutex>
#include <sstream>
#include <tuple>

#include "lower.h"

namespace tsb {
void contract2_impl(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires, int dl, const void* rhs,
                    int rhs_rank, const int32_t* rhs_wires, int dr, void* out, int out_rank,
                    const int32_t* out_wires, int dout, int n_wires, const int32_t* wext, int mode, bool accumulate,
                    bool force_generic, cudaStream_t stream);
void rewire_impl(tsb_ctx* ctx, const void* src, int src_rank, const int32_t* sext, const int32_t* fixed_idx, int ds,
                 void* dst, int out_rank, const int32_t* out_ports, int dd, bool clamp, bool accumulate,
                 cudaStream_t stream);
struct GemmShape {
  bool ok = false;
  bool swap = false;
  BitLabels B, M, N, K;
};
GemmShape gemm_shape(const BitLabels& lb, const BitLabels& rb, const BitLabels& ob, int out_elem_size, double min_intensity,
                     bool boolean_operands);
GenDesc lower_generic_contract(int lhs_rank, const int32_t* lhs_wires, int rhs_rank, const int32_t* rhs_wires,
                               int out_rank, const int32_t* out_wires, const int32_t* wext, int dl, int dr, int dout,
                               bool clamp, bool accumulate, bool& out_dups);

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
  bool post = false;             // shared node computed AFTER the slice loop (tail of the tree)
  bool gathered = false;         // per-slice node stored with the sliced wires leading: slice s fills block s
  bool loop_local = false;       // lives only inside one slice iteration (a second slice lane works on a copy)
  bool packed = false;           // boolean relation stored one bit per entry (k_merge_bits.cuh)
  int64_t stored_numel() const { return reg_numel * ((shared || gathered) ? s_numel : 1); }
  std::vector<int32_t> stored() const {  // axis order of the stored buffer
    std::vector<int32_t> v;
    if (shared || gathered) v = slab;
    v.insert(v.end(), reg.begin(), reg.end());
    return v;
  }
};

enum OpKind { OP_LEAF = 0, OP_STEP = 1, OP_FINAL_PERMUTE = 2, OP_PACK = 3, OP_UNPACK = 4, OP_EXPAND = 5 };

struct POp {
  int kind = OP_STEP;
  bool per_slice = false;
  int lhs = -1, rhs = -1, out = -1;  // node ids
  bool use_merge = false;
  bool use_gemm = false;  // lowered at run time: transposes + tcgen05 int8 GEMM
  bool use_bits = false;  // both operands and the result are bit-packed: k_merge_bits
  double alg_bytes_u8 = 0;  // algorithmic bytes at one byte per boolean entry
  std::vector<int> xn;                    // fused step: operands beyond lhs / rhs (node ids) ...
  std::vector<std::vector<int32_t>> xw;   // ... and their wires
  int n_fused = 0;                        // tree steps folded into this launch besides its own
  BitsNDesc nd{};                         // descriptor of a fused step (k_merge_bits_n)
  BitsFDesc fd{};                         // ... or of its factored form (k_merge_bits_f), when use_f
  long wf_chosen = 0, wf_default = 0;     // modelled L1 wavefronts per work item (make_bits_f_desc)
  double launch_entries = 0;              // fused step: entries the launch must read and write (operands + result)
  bool use_f = false;
  int64_t tab_off = -1;                   // use_f: offset (uint32 units) of this step's offset tables in the plan's d_tabs
  int n_operands() const { return kind == OP_STEP ? 2 + (int)xn.size() : 1; }
  std::vector<int> all_operands() const { std::vector<int> v{lhs, rhs}; v.insert(v.end(), xn.begin(), xn.end()); return v; }
  int operand(int j) const { return j == 0 ? lhs : j == 1 ? rhs : xn[j - 2]; }
  const std::vector<int32_t>& operand_wires(int j) const { return j == 0 ? lw : j == 1 ? rw : xw[j - 2]; }
  std::vector<int32_t>& operand_wires(int j) { return j == 0 ? lw : j == 1 ? rw : xw[j - 2]; }
  bool cv_map = false;      // conversion with an index map (broadcast over missing wires / packed source)
  BitRuns cv_lo{}, cv_hi{}; // destination lane (bits 0..4) / word index -> source entry offset
  MergeDesc md{};
  // generic description (wires + extents) used when the merge kernel does not apply
  std::vector<int32_t> lw, rw, ow;
  // leaf preparation
  int leaf = -1;
  std::vector<int32_t> leaf_ports;  // dst axis j <- leaf axis leaf_ports[j]
  double alg_bytes = 0, alg_ops = 0;
  int b = 0, m = 0, n = 0, k = 0;  // label class sizes in bits (for reports)
  float last_ms = 0;
  // generic-kernel descriptor, lowered once at plan time
  bool has_gd = false, gd_zero = false;
  GenDesc gd{};
  bool small = false;  // runs inside a k_small_batch launch
  int wave = 0;        // dependency depth among the once-only ops
};

struct Launch {   // one entry of the once-only schedule
  int wave = 0;
  int op = -1;          // single op, or
  int batch_begin = 0;  // [begin, end) into the small-op array
  int batch_end = 0;
  bool chain = false;   // the small ops run in order inside one CTA (k_small_chain)
  int split = 1;        // CTAs per op of a batch (waves that hold the larger small ops)
  float last_ms = 0;
};

}  // namespace tsb

using namespace tsb;

struct tsb_plan {
  tsb_ctx* ctx = nullptr;
  // slice-range restriction of the once-only ("pre") ops during one execution: the first rr_nfixed sliced wires
  // are fixed to rr_digits (the range is an aligned power-of-two block of slice ids), see pre_block()
  int rr_nfixed = 0;
  std::vector<int64_t> rr_digits;
  int rr_launches = 0;  // sub-block launches of the last execution (reported by tsb_plan_describe)
  // side streams for the mutually independent launches of one wave of the once-only schedule (fork / join)
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> side_ev;   // [0] fork, [1 + i] join of side stream i
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
  int timing = 0;                      // 0: off; 1: an event pair + host sync around every launch (tools); 2: four
                                       // events at the phase boundaries of an execution, no host sync (bench.py)
  cudaEvent_t phase_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // start | pre done | slice loop done | end
  bool phase_valid = false;
  std::vector<cudaEvent_t> ev;
  std::vector<Launch> schedule;        // once-only part before the slice loop, in launch order
  std::vector<Launch> post_schedule;   // once-only part after the slice loop (re-batched tail of the tree)
  std::vector<int> gathered_nodes;     // zero-filled before the slice loop
  std::vector<int> slice_ops;          // per-slice ops, in order
  int64_t lane_shift = 0;              // > 0: two slice lanes; the loop-local nodes of the second lane live this far up
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t lane_ev[2] = {nullptr, nullptr};
  std::vector<SmallOp> small_ops;      // host copy
  SmallOp* d_small_ops = nullptr;
  uint32_t* d_tabs = nullptr;          // offset tables of the factored fused steps, computed once (bits_f_tables)
  uint8_t* hx_leaves = nullptr;        // tsb_plan_execute_host: device staging of the leaves (back to back) ...
  void* hx_out = nullptr;              // ... and of the result
  size_t hx_total = 0;
  std::vector<const void*> hx_ptrs;    // leaf pointers into hx_leaves
  std::mutex hx_mu;                    // one host-buffer execution at a time (they share the staging buffers)
  void** d_ptrs = nullptr;             // device pointer table: [arena, out, leaf 0, leaf 1, ...]
  std::vector<void*> h_ptrs;           // host copy of the table (pageable: the async copy stages it)
  bool mask_root = false;              // root keeps the sliced wires; a final masked sum selects the executed slices
  int64_t mask_offset = -1;            // arena offset of the slice mask (u8, one entry per slice combination)
  std::vector<void*> h_ptrs_uploaded;  // what d_ptrs holds (the upload is skipped when nothing changed)
  // ---- executions are serialised: one at a time per plan, and one on another stream waits for the previous one
  std::mutex mu;
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  cudaEvent_t done_ev = nullptr;
  // ---- CUDA-graph replay: the launch sequence of an execution is static per (slice range, result
  // buffer); from the second execution with the same key on, it is captured once and replayed with one host call
  struct GraphKey {
    void* out; int64_t begin, end; uint64_t leaf_hash;
    bool operator<(const GraphKey& o) const {
      return std::tie(out, begin, end, leaf_hash) < std::tie(o.out, o.begin, o.end, o.leaf_hash);
    }
  };
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    tsb::HeldBlocks held;                      // temporaries of the captured launches (owned until the graph dies)
    std::map<std::string, int64_t> kcount;     // kernels per replay
    int64_t launches = 0;
    int rr_launches = 0;
    uint64_t last_used = 0;
  };
  std::map<GraphKey, GraphEntry> graphs;
  std::map<GraphKey, int> seen;          // executions per key so far (capture on the second)
  std::set<GraphKey> ungraphable;        // capture failed once: always direct
  cudaStream_t cap_stream = nullptr;     // capture happens on a stream of the plan's own (the caller's may be the
                                         // legacy default stream, which cannot be captured); replays go to the caller's
  void* gblock = nullptr;                // arena block owned by the graphs of this plan (same addresses on replay)
  void* gres = nullptr;                  // small results land here inside a graph, then one D2D copy to the caller
  uint64_t tick = 0;
  int64_t n_replays = 0, n_captures = 0;
  bool bakes_leaf_ptrs = false;          // some launch takes a leaf pointer as a kernel argument (not via d_ptrs)
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
  // once-only ("pre" / "post") limit: 16 MiB per tensor with all slices kept, in either representation
  const bool will_pack = P.mode == TSB_SAT && !(P.flags & (TSB_PLAN_FORCE_GENERIC | TSB_PLAN_NO_PACK));
  const int batch_log2 = D.batch_log2 >= 0 ? D.batch_log2 : (will_pack ? 27 : 24);
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
  // sliced output wires: per-slice results are stacked into an intermediate buffer
  for (auto w : P.out_wires) if (S.count(w)) P.stack_out = true;
  // Tail of the tree: once the per-slice tensors are small again (with the sliced wires kept), the slices
  // are gathered into one tensor and the remaining steps run ONCE after the slice loop.
  if (sliced_any && n_steps > 0 && !P.stack_out && prod_ext(P.sliced, P.wext) <= 4096) {
    std::vector<int> parent(2 * L - 1, -1);
    for (int s = 0; s < n_steps; s++) { parent[path[s].first] = L + s; parent[path[s].second] = L + s; }
    auto lg_b = [&](int i) { return std::log2((double)P.nodes[i].reg_numel) + std::log2((double)P.nodes[i].s_numel); };
    for (int i = 2 * L - 2; i >= L; i--) {
      PNode& nd = P.nodes[i];
      if (nd.shared) continue;
      const bool parent_post = i == P.root || P.nodes[parent[i]].post;
      nd.post = parent_post && lg_b(i) <= batch_log2;
    }
    // a post node may not gather a child that is huge once all its slices are kept
    for (int i = L; i <= 2 * L - 2; i++) {
      PNode& nd = P.nodes[i];
      if (!nd.post) continue;
      for (int c : {path[i - L].first, path[i - L].second})
        if (!P.nodes[c].shared && !P.nodes[c].post && lg_b(c) > batch_log2 + 6) nd.post = false;
    }
    // post status must be upward closed: a node whose parent lost it keeps it only if the parent still has it
    for (int i = 2 * L - 2; i >= L; i--)
      if (P.nodes[i].post && i != P.root && !P.nodes[parent[i]].post) {
        // parent runs per slice: this node would have to be per slice as well
        P.nodes[i].post = false;
      }
    for (int i = L; i <= 2 * L - 2; i++) {
      PNode& nd = P.nodes[i];
      if (nd.post) { nd.shared = true; continue; }
      if (!nd.shared && i != P.root && P.nodes[parent[i]].post) nd.gathered = true;
    }
    // A post root keeps the sliced wires as leading axes; one final masked sum over them selects the slices
    // this call was asked to execute (other ranks execute the others), so that contributions that never
    // pass through the slice loop are not counted once per rank.
    P.mask_root = P.nodes[P.root].post;
  }
  if (P.mode == TSB_COUNT) {
    P.nodes[P.root].dtype = TSB_I64;
    double tot = bound[P.root];
    for (auto w : P.sliced) if (freew(w)) tot += std::log2((double)P.wext[w]);
    P.info.may_overflow = tot > 62.0;
    for (int i = 0; i < 2 * L - 1; i++) if (bound[i] > 62.0) P.info.may_overflow = 1;
  }
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
    op.alg_bytes_u8 = op.alg_bytes;
    op.alg_ops = 2.0 * std::exp2(lb + lm + ln + lk);
    P.ops.push_back(op);
  }
  if (n_steps > 0 && !P.stack_out && !P.mask_root) P.nodes[P.root].is_out = true;

  // ---- kernel selection ------------------------------------------------------------------------
  const bool force_generic = (P.flags & TSB_PLAN_FORCE_GENERIC) != 0;
  const bool clamp = P.mode == TSB_SAT;
  for (auto& op : P.ops) {
    const PNode& o = P.nodes[op.out];
    const int64_t out_numel = prod_ext(op.kind == OP_LEAF ? o.stored() : op.ow, P.wext);
    if (op.kind == OP_STEP && !force_generic && out_numel > 4096) {
      bool pow2 = true;
      for (auto w : op.lw) pow2 = pow2 && is_pow2(P.wext[w]);
      for (auto w : op.rw) pow2 = pow2 && is_pow2(P.wext[w]);
      if (pow2) {
        BitLabels lb = bit_labels(op.lw.data(), (int)op.lw.size(), P.wext.data());
        BitLabels rb = bit_labels(op.rw.data(), (int)op.rw.size(), P.wext.data());
        BitLabels ob = bit_labels(op.ow.data(), (int)op.ow.size(), P.wext.data());
        bool to_gemm = false;
#ifndef TSB_HOST_EMU
        // steps above the ridge are lowered at run time by contract2_impl (transpose -> tcgen05 GEMM -> transpose)
        if (!(P.flags & TSB_PLAN_NO_GEMM) && P.nodes[op.lhs].dtype == TSB_U8 && P.nodes[op.rhs].dtype == TSB_U8 &&
            (clamp ? o.dtype == TSB_U8 : o.dtype != TSB_U8) && !(op.per_slice && op.out == P.root))
          to_gemm = gemm_shape(lb, rb, ob, dtype_size(o.dtype), kGemmMinIntensity, clamp).ok;
#endif
        MergeArgs m = make_merge_desc(lb, &rb, ob, false);
        if (m.ok && m.d.kbits > kSplitKBits) to_gemm = true;  // run-time split-K
        if (to_gemm) {
          op.use_gemm = true;
        } else if (m.ok) {
          op.use_merge = true;
          op.md = m.d;
        }
      }
    }
    if (op.kind == OP_STEP && !force_generic && !op.use_merge && !op.use_gemm) {
      // many summed wires but too few outputs for the merge kernel: run-time lowering splits the sum
      std::set<int32_t> so(op.ow.begin(), op.ow.end());
      double kb = 0;
      bool pow2 = true;
      for (auto w : op.lw) { pow2 = pow2 && is_pow2(P.wext[w]); if (!so.count(w)) kb += std::log2((double)P.wext[w]); }
      for (auto w : op.rw) pow2 = pow2 && is_pow2(P.wext[w]);
      if (pow2 && kb > 12) op.use_gemm = true;
    }
    if (op.use_merge || op.use_gemm) continue;
    // generic kernel: lowered once, here (the root of a sliced plan accumulates and stays on the slow path)
    if (op.kind == OP_LEAF) {
      const auto& lw = P.leaf_wires[op.leaf];
      const std::vector<int32_t> ow = o.stored();
      op.gd = lower_generic_contract((int)lw.size(), lw.data(), 0, nullptr, (int)ow.size(), ow.data(), P.wext.data(),
                                     TSB_U8, TSB_U8, o.dtype, clamp, false, op.gd_zero);
      op.gd.has_rhs = 0;
      op.has_gd = true;
    } else if (!(op.per_slice && op.out == P.root)) {
      op.gd = lower_generic_contract((int)op.lw.size(), op.lw.data(), (int)op.rw.size(), op.rw.data(),
                                     (int)op.ow.size(), op.ow.data(), P.wext.data(), P.nodes[op.lhs].dtype,
                                     P.nodes[op.rhs].dtype, o.dtype, clamp, false, op.gd_zero);
      op.has_gd = true;
    }
    if (op.has_gd && !op.per_slice && !(P.n_leaves == 1 && sliced_any)) {
      const GenDesc& g = op.gd;
      bool ok = g.n_free + g.n_red <= kSmallMaxDims && g.n_out_iter <= (1 << 14) && g.n_red_iter <= (1 << 12) &&
                g.n_out_iter * g.n_red_iter <= (1 << 17);  // one CTA of 64 threads: keep it in the microsecond range
      for (int q = 0; ok && q < g.n_free + g.n_red; q++)
        ok = g.sl[q] < (int64_t(1) << 30) && g.sr[q] < (int64_t(1) << 30) && g.so[q] < (int64_t(1) << 30);
      op.small = ok;
    }
  }

  // ---- bit-packed representation for the boolean semiring -----------------------------
  // A step that touches a tensor of >= 2^kPackMinLog2 entries runs on bit-packed operands (k_merge_bits):
  // 8x fewer bytes on the steps that are HBM-bound.  Conversions are inserted where a tensor crosses between
  // the two representations; by construction they fall on small tensors (the leaves' side of the tree and the
  // gathered tail), except around GEMM-lowered steps.  A small operand is additionally stored BROADCAST over
  // those of the consumer's five lowest output wires it does not depend on, so that inside the hot kernel
  // every operand word is a plain load (no in-word lane expansion).  Consecutive packed steps are FUSED into
  // one n-ary launch (k_merge_bits_n) when the intermediate between them is large: it is then neither
  // written nor read back (bucket elimination without materialising the product).
  if (P.mode == TSB_SAT && !force_generic && !(P.flags & TSB_PLAN_NO_PACK)) {
    // test knobs: thresholds lowered so that small parity cases take these paths
    auto env_or = [](const char* name, double dflt) { const char* e = std::getenv(name); return e ? std::atof(e) : dflt; };
    const double kPackMinLog2 = env_or("TSB_PACK_MIN_LOG2", 20.0);
    const double kExpandMaxLog2 = env_or("TSB_PACK_EXPAND_LOG2", 27.0);  // broadcast twins stay L2-sized
    const double kFuseMinLog2 = env_or("TSB_FUSE_MIN_LOG2", 18.0);       // fuse across intermediates at least this big (22 -> 18: the small per-slice chain of the n=200 plan, 56 -> 32 launches per slice)
    const int kFuseMaxK = (P.flags & TSB_PLAN_NO_FUSE) ? -1 : (int)env_or("TSB_FUSE_MAX_K", 2.0);
    const int kFusePair = (int)env_or("TSB_FUSE_PAIR", 3.0);  // pair wires of the factored kernel (2, 3: private streaming); -1: k_merge_bits_n
    const int NN0 = (int)P.nodes.size();
    auto bits_desc = [&](POp& op) {
      BitLabels lb = bit_labels(op.lw.data(), (int)op.lw.size(), P.wext.data());
      BitLabels rb = bit_labels(op.rw.data(), (int)op.rw.size(), P.wext.data());
      BitLabels ob = bit_labels(op.ow.data(), (int)op.ow.size(), P.wext.data());
      return make_merge_desc(lb, &rb, ob, false, 7);
    };
    for (auto& op : P.ops) {
      if (op.kind != OP_STEP || !op.use_merge) continue;
      if (P.stack_out && op.out == P.root) continue;
      const double big = std::max({(double)prod_ext(op.lw, P.wext), (double)prod_ext(op.rw, P.wext),
                                   (double)prod_ext(op.ow, P.wext)});
      if (std::log2(big) < kPackMinLog2) continue;
      MergeArgs m = bits_desc(op);
      if (!m.ok || m.d.kbits > 8) continue;
      op.use_bits = true;
      op.md = m.d;
    }
    auto has_role = [&](int i) { return P.nodes[i].is_out || P.nodes[i].gathered || i == P.root; };
    // wires of the n_low lowest output positions of `ow` that the operand does not hold (5: the in-word lanes,
    // enough for the two-operand kernels; 7: the whole 16-byte vector, what a fused step needs)
    auto missing_low = [&](const std::vector<int32_t>& ow, const std::vector<int32_t>& w, int n_low) {
      std::vector<int32_t> add;
      for (int q = std::max(0, (int)ow.size() - n_low); q < (int)ow.size(); q++)
        if (std::find(w.begin(), w.end(), ow[q]) == w.end()) add.push_back(ow[q]);
      return add;
    };
    // the slice of a broadcast twin a step reads must stay L2-sized; the whole twin (all slices) is written once
    auto can_broadcast = [&](int node, size_t n_add) {
      return n_add == 0 || (std::log2((double)P.nodes[node].reg_numel) + (double)n_add <= kExpandMaxLog2 &&
                            std::log2((double)P.nodes[node].stored_numel()) + (double)n_add <= kExpandMaxLog2 + 4.0);
    };
    // ---- fusion ------------------------------------------------------------------------------------------
    std::vector<char> dead(P.ops.size(), 0);
    if (kFuseMaxK >= 0) {
      // candidate edges producer -> consumer, largest intermediate first: the widest tensors are the ones
      // worth never storing; a fused group is represented by its topmost consumer
      std::vector<int> producer_op(NN0, -1), absorbed(P.ops.size(), -1);
      for (int oi = 0; oi < (int)P.ops.size(); oi++) producer_op[P.ops[oi].out] = oi;
      struct Edge { double lg; int pi, oi; };
      std::vector<Edge> edges;
      for (int oi = 0; oi < (int)P.ops.size(); oi++) {
        const POp& op = P.ops[oi];
        if (!op.use_bits) continue;
        for (int j = 0; j < 2; j++) {
          const int T = op.operand(j);
          const int pi = producer_op[T];
          if (pi < 0 || !P.ops[pi].use_bits || has_role(T)) continue;
          const POp& s1 = P.ops[pi];
          if (s1.per_slice != op.per_slice || P.nodes[s1.out].post != P.nodes[op.out].post) continue;
          if (op.operand_wires(j) != s1.ow) continue;
          const double lg = std::log2((double)prod_ext(s1.ow, P.wext));
          if (lg >= kFuseMinLog2) edges.push_back({lg, pi, oi});
        }
      }
      std::stable_sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return a.lg > b.lg; });
      auto find_root = [&](int oi) { while (absorbed[oi] >= 0) oi = absorbed[oi]; return oi; };
      for (const Edge& e : edges) {
        const int ri = find_root(e.oi);
        POp& op = P.ops[ri];
        const POp& s1 = P.ops[e.pi];  // still the root of its own group: its output edge is this one
        const int T = s1.out;
        const int n_cur = op.n_operands(), n_s1 = s1.n_operands();
        int j = -1;
        for (int q = 0; q < n_cur; q++) if (op.operand(q) == T) j = q;
        if (j < 0 || n_cur - 1 + n_s1 > kBitsMaxOps) continue;
        std::vector<int> nodes;
        std::vector<std::vector<int32_t>> wires;
        for (int q = 0; q < n_cur; q++)
          if (q != j) { nodes.push_back(op.operand(q)); wires.push_back(op.operand_wires(q)); }
        for (int q = 0; q < n_s1; q++) { nodes.push_back(s1.operand(q)); wires.push_back(s1.operand_wires(q)); }
        std::set<int32_t> summed;
        bool ok = true;
        for (size_t q = 0; q < nodes.size() && ok; q++) {
          for (auto w : wires[q])
            if (std::find(op.ow.begin(), op.ow.end(), w) == op.ow.end()) summed.insert(w);
          ok = can_broadcast(nodes[q], missing_low(op.ow, wires[q], 7).size()) && wires[q].size() <= 36;
        }
        if (!ok || (int)summed.size() > std::min(kFuseMaxK, (int)kBitsMaxKB)) continue;
        // fuse: the producer disappears, its operands join the consumer's launch
        op.lhs = nodes[0]; op.lw = wires[0];
        op.rhs = nodes[1]; op.rw = wires[1];
        op.xn.assign(nodes.begin() + 2, nodes.end());
        op.xw.assign(wires.begin() + 2, wires.end());
        op.launch_entries = (double)prod_ext(op.ow, P.wext);
        for (auto& w : wires) op.launch_entries += (double)prod_ext(w, P.wext);
        op.alg_bytes += s1.alg_bytes;
        op.alg_bytes_u8 += s1.alg_bytes_u8;
        op.alg_ops += s1.alg_ops;
        op.n_fused += 1 + s1.n_fused;
        dead[e.pi] = 1;
        absorbed[e.pi] = ri;
      }
    }
    std::vector<char> made_packed(NN0, 0);
    std::vector<int> consumer(NN0, -1);
    for (int oi = 0; oi < (int)P.ops.size(); oi++) {
      if (dead[oi]) continue;
      const POp& op = P.ops[oi];
      if (op.use_bits) made_packed[op.out] = 1;
      for (int j = 0; j < op.n_operands(); j++)
        if (op.operand(j) >= 0) consumer[op.operand(j)] = oi;
    }
    std::vector<int> prod_rep(NN0), cons_rep(NN0);
    std::vector<std::vector<POp>> after(NN0);  // conversions emitted right after the producer of a node
    for (int i = 0; i < NN0; i++) prod_rep[i] = cons_rep[i] = i;
    for (int i = 0; i < NN0; i++) {
      bool role = has_role(i);  // bytes must exist in node i
      const bool Pp = made_packed[i] != 0;
      const bool Cp = consumer[i] >= 0 && P.ops[consumer[i]].use_bits;
      if (!Pp && !Cp) continue;
      // a gathered tensor between two packed steps stays packed: slice s fills bit block s (16-byte aligned when a
      // block holds at least 128 entries), the tail of the tree reads it as it is
      static const bool pack_gathered = [] { const char* e = getenv("TSB_PACK_GATHERED"); return !(e && e[0] == '0'); }();
      if (pack_gathered && role && P.nodes[i].gathered && !P.nodes[i].is_out && i != P.root && Pp && Cp &&
          P.nodes[i].reg_numel % 128 == 0)
        role = false;
      const bool bytes_in_i = !Pp || role || (consumer[i] >= 0 && !Cp);
      if (Pp) {
        if (bytes_in_i) {  // the step writes a packed temporary that is unpacked into node i
          PNode t = P.nodes[i];
          t.packed = true; t.is_out = false; t.gathered = false; t.offset = -1;
          prod_rep[i] = (int)P.nodes.size();
          P.nodes.push_back(t);
          POp cv;
          cv.kind = OP_UNPACK; cv.lhs = prod_rep[i]; cv.out = i;
          cv.per_slice = !P.nodes[i].shared;
          after[i].push_back(cv);
        } else {
          P.nodes[i].packed = true;
        }
      }
      if (!Cp) continue;
      POp& z = P.ops[consumer[i]];
      int zj = 0;
      for (int j = 0; j < z.n_operands(); j++) if (z.operand(j) == i) zj = j;
      std::vector<int32_t> add = missing_low(z.ow, z.operand_wires(zj), z.xn.empty() ? 5 : 7);
      const PNode src = P.nodes[i];
      if (!can_broadcast(i, add.size())) add.clear();
      if (add.empty() && src.packed) continue;  // read as is
      PNode t = src;
      t.packed = true; t.is_out = false; t.offset = -1;
      if (t.gathered) { t.gathered = false; t.shared = true; t.post = true; }  // converted once, after the slice loop
      {
        std::set<int32_t> rs(t.reg.begin(), t.reg.end());
        rs.insert(add.begin(), add.end());
        t.reg = order_reg(rs);
        t.reg_numel = prod_ext(t.reg, P.wext);
      }
      cons_rep[i] = (int)P.nodes.size();
      P.nodes.push_back(t);
      POp cv;
      cv.kind = src.packed ? OP_EXPAND : OP_PACK;
      cv.lhs = i; cv.out = cons_rep[i];
      cv.per_slice = !t.shared;
      const std::vector<int32_t> dw = t.stored(), sw = src.stored();
      BitLabels db = bit_labels(dw.data(), (int)dw.size(), P.wext.data());
      BitLabels sb = bit_labels(sw.data(), (int)sw.size(), P.wext.data());
      if (src.packed && db.size() >= 7) {
        // packed -> packed broadcast = a one-operand merge whose missing output wires repeat the value
        MergeArgs bm = make_merge_desc(sb, nullptr, db, false, 7, true);
        TSB_CHECK(bm.ok && bm.d.kbits == 0, TSB_ERR_INVALID, "internal: broadcast twin is not a monotone merge");
        cv.md = bm.d;
        cv.use_bits = true;
      } else if (!add.empty() || src.packed) {
        // dst entry index -> src entry index (bits 0..4 of dst: lanes of one word; the rest: word index)
        auto sp = positions(sb);
        const int nd = (int)db.size();
        for (int j = 0; j < nd; j++) {
          auto it = sp.find(db[nd - 1 - j]);
          if (it == sp.end()) continue;
          if (j < 5) runs_add(cv.cv_lo, j, it->second);
          else runs_add(cv.cv_hi, j - 5, it->second);
        }
        cv.cv_map = true;
      }
      after[i].push_back(cv);
      // the consumer sees the broadcast operand
      z.operand_wires(zj) = P.nodes[z.out].shared ? t.stored() : t.reg;
    }
    std::vector<POp> ops2;
    ops2.reserve(P.ops.size() + 64);
    for (int oi = 0; oi < (int)P.ops.size(); oi++) {
      if (dead[oi]) continue;
      POp op = P.ops[oi];
      const int y = op.out;
      if (op.use_bits) {
        op.lhs = cons_rep[op.lhs];
        op.rhs = cons_rep[op.rhs];
        for (auto& x : op.xn) x = cons_rep[x];
        op.out = prod_rep[y];
        if (op.xn.empty()) {
          MergeArgs m = bits_desc(op);  // operand wire lists may have grown
          TSB_CHECK(m.ok && m.d.kbits <= 8, TSB_ERR_INVALID, "internal: broadcast operand broke the merge order");
          op.md = m.d;
        } else {
          std::vector<BitLabels> ol;
          for (int j = 0; j < op.n_operands(); j++) {
            const auto& w = op.operand_wires(j);
            ol.push_back(bit_labels(w.data(), (int)w.size(), P.wext.data()));
          }
          const BitLabels ob = bit_labels(op.ow.data(), (int)op.ow.size(), P.wext.data());
          BitsNArgs m = make_bits_n_desc(ol, ob);
          TSB_CHECK(m.ok, TSB_ERR_INVALID, "internal: fused step is not a word-aligned monotone merge");
          op.nd = m.d;
          if (kFusePair >= 0 && m.d.kbits <= kBitsFMaxKB) {
            int n_slab_bits = 0;
            if (P.nodes[y].shared)
              for (auto w : P.nodes[y].slab) n_slab_bits += ilog2(P.wext[w]);
            BitsFArgs f = make_bits_f_desc(ol, ob, kFusePair, n_slab_bits, env_or("TSB_FUSE_LANES", 0.0) != 0.0,
                                           (int)env_or("TSB_FUSE_TILE", 0.0));
            if (!f.ok && env_or("TSB_FUSE_TILE", 0.0) != 0.0)   // too many bit runs for the tables: plain order
              f = make_bits_f_desc(ol, ob, kFusePair, n_slab_bits, false, false);
            if (f.ok) { op.fd = f.d; op.use_f = true; op.wf_chosen = f.wf_chosen; op.wf_default = f.wf_default; }
          }
          P.info.n_fused_steps += op.n_fused;
        }
        // bytes this launch must move at one bit per entry; a fused launch neither writes nor re-reads the
        // intermediates of the steps it folds in (alg_bytes_u8 keeps SURVEY 8(d)'s per-step sum at one byte per entry)
        op.alg_bytes = (op.launch_entries > 0 ? op.launch_entries : op.alg_bytes_u8) / 8.0;
        P.info.n_packed_steps += 1 + op.n_fused;
      }
      ops2.push_back(op);
      for (auto& cv : after[y]) { ops2.push_back(cv); P.info.n_conversions++; }
    }
    P.ops.swap(ops2);
  }
  const int NN = (int)P.nodes.size();  // tree nodes + packed twins (the stacked result buffer comes later)

  // ---- schedule of the once-only parts: waves of mutually independent ops ---------------------------
  // pre = before the slice loop, post = after it (tail of the tree on gathered slices)
  std::vector<int> producer(P.nodes.size() + 1, -1);
  for (int i = 0; i < (int)P.ops.size(); i++) producer[P.ops[i].out] = i;
  auto is_post_op = [&](const POp& op) { return !op.per_slice && P.nodes[op.out].post; };
  for (auto& op : P.ops) {
    if (op.per_slice) continue;
    const bool post = is_post_op(op);
    int w = 0;
    for (int opnd : op.all_operands())
      if (opnd >= 0 && producer[opnd] >= 0) {
        const POp& pr = P.ops[producer[opnd]];
        if (!pr.per_slice && is_post_op(pr) == post) w = std::max(w, pr.wave + 1);
      }
    op.wave = w;
  }
  // per-slice ops: dependency level inside one slice iteration (ops of one level are mutually independent)
  for (auto& op : P.ops) {
    if (!op.per_slice) continue;
    int w = 0;
    for (int opnd : op.all_operands())
      if (opnd >= 0 && producer[opnd] >= 0 && P.ops[producer[opnd]].per_slice) w = std::max(w, P.ops[producer[opnd]].wave + 1);
    op.wave = w;
  }
  std::vector<int> pre_ops, post_ops;
  for (int i = 0; i < (int)P.ops.size(); i++) {
    if (P.ops[i].per_slice) P.slice_ops.push_back(i);
    else (is_post_op(P.ops[i]) ? post_ops : pre_ops).push_back(i);
  }
  auto by_wave = [&](int a, int b) { return P.ops[a].wave < P.ops[b].wave; };
  std::stable_sort(pre_ops.begin(), pre_ops.end(), by_wave);
  std::stable_sort(post_ops.begin(), post_ops.end(), by_wave);

  // ---- memory plan -----------------------------------------------------------------------------
  // timeline: [pre waves][per-slice ops, repeated for every slice][post waves]; ops of one wave run
  // concurrently.  Whatever must survive from one slice iteration to the next (shared operands of the
  // loop, gathered results) is allocated outside the loop's reusable region.
  for (int i = 0; i < (int)P.nodes.size(); i++) {
    PNode& nd = P.nodes[i];
    nd.bytes = nd.packed ? packed_bytes(nd.stored_numel()) : (size_t)(nd.stored_numel() * dtype_size(nd.dtype));
  }
  const int n_pre = pre_ops.empty() ? 0 : P.ops[pre_ops.back()].wave + 1;
  const int n_loop = (int)P.slice_ops.size();
  const int n_post = post_ops.empty() ? 0 : P.ops[post_ops.back()].wave + 1;
  std::vector<int> tpos(P.ops.size(), 0);
  for (int i : pre_ops) tpos[i] = P.ops[i].wave;
  for (int q = 0; q < n_loop; q++) tpos[P.slice_ops[q]] = n_pre + q;
  for (int i : post_ops) tpos[i] = n_pre + n_loop + P.ops[i].wave;
  const int T_END = n_pre + n_loop + n_post;
  std::vector<int> last_use(P.nodes.size(), -1);
  for (int i = 0; i < (int)P.ops.size(); i++)
    for (int opnd : P.ops[i].all_operands())
      if (opnd >= 0) last_use[opnd] = std::max(last_use[opnd], tpos[i]);
  // shared tensors consumed inside the loop must survive every slice iteration
  for (int i : P.slice_ops)
    for (int opnd : P.ops[i].all_operands())
      if (opnd >= 0 && P.nodes[opnd].shared) last_use[opnd] = std::max(last_use[opnd], n_pre + n_loop - 1);
  FreeList fl;
  if (P.mask_root) P.mask_offset = fl.alloc((int64_t)P.nodes[P.root].s_numel);
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
  {
    std::vector<std::vector<int>> at(T_END + 1);
    for (int i = 0; i < (int)P.ops.size(); i++) at[tpos[i]].push_back(i);
    for (int t = 0; t < T_END; t++) {
      if (t == n_pre) {
        for (int i = 0; i < NN; i++)
          if (P.nodes[i].gathered) {
            P.nodes[i].offset = fl.alloc((int64_t)P.nodes[i].bytes);
            P.gathered_nodes.push_back(i);
          }
      }
      for (int i : at[t]) {
        PNode& o = P.nodes[P.ops[i].out];
        if (!o.is_out && !o.gathered) o.offset = fl.alloc((int64_t)o.bytes);
      }
      std::set<int> freed;
      for (int i : at[t])
        for (int opnd : P.ops[i].all_operands())
          if (opnd >= 0 && last_use[opnd] == t && P.nodes[opnd].offset >= 0 && !freed.count(opnd)) {
            fl.release(P.nodes[opnd].offset, (int64_t)P.nodes[opnd].bytes);
            freed.insert(opnd);
          }
      if (t == n_pre + n_loop - 1) {
        // end of the loop: shared operands whose last consumer was inside the loop, at an earlier position
        for (int nd = 0; nd < NN; nd++)
          if (P.nodes[nd].shared && P.nodes[nd].offset >= 0 && last_use[nd] == t && !freed.count(nd)) {
            bool used_here = false;
            for (int i : at[t])
              for (int opnd : P.ops[i].all_operands()) used_here = used_here || opnd == nd;
            if (!used_here) { fl.release(P.nodes[nd].offset, (int64_t)P.nodes[nd].bytes); freed.insert(nd); }
          }
      }
    }
  }
  P.info.peak_bytes = fl.peak;
  {
    // Two slice lanes (TSB_SLICE_LANES=2; under graph capture the slices alternate between two parallel branches, so
    // that the launches of one slice fill the SMs the other leaves idle at its kernel boundaries): every node that
    // lives only inside one slice iteration gets a second copy above the block.  Only for plans whose per-slice ops
    // need nothing but their descriptors and the block (no temporaries, no accumulation into the result).
    // Default (TSB_SLICE_LANES=0): only boolean plans whose slice is a long chain of launches (>= 12: the
    // n=200 plan has 41, mostly small ones — sub-cube 63.6 -> 54.8 ms); the n=50 plan's slice is four wide launches
    // that already fill the GPU, and a second copy of its intermediates in flight costs 4 % (L2).  1: never, 2: always.
    static const int lanes_env = [] { const char* e = getenv("TSB_SLICE_LANES"); return e ? atoi(e) : 0; }();
    int64_t n_sl = 1;
    for (auto w : P.sliced) n_sl *= P.wext[w];
    bool all_bits = P.mode == TSB_SAT;
    for (int i : P.slice_ops) all_bits = all_bits && (P.ops[i].kind != OP_STEP || P.ops[i].use_bits || P.ops[i].use_merge);
    const bool want = lanes_env >= 2 || (lanes_env == 0 && all_bits && P.slice_ops.size() >= 12);
    bool ok = want && sliced_any && L > 1 && !P.stack_out && n_sl >= 2 && !P.slice_ops.empty();
    int64_t lo = INT64_MAX, hi = 0;
    for (int i : P.slice_ops) {
      const POp& op = P.ops[i];
      const PNode& o = P.nodes[op.out];
      const bool plain = op.kind == OP_PACK || op.kind == OP_UNPACK || op.kind == OP_EXPAND ||
                         (op.kind == OP_STEP && !op.use_gemm && (op.use_bits || op.use_merge));
      ok = ok && plain && !o.is_out && op.out != P.root;
      if (!o.gathered && !o.is_out && o.offset >= 0) {
        lo = std::min<int64_t>(lo, o.offset);
        hi = std::max<int64_t>(hi, o.offset + (int64_t)o.bytes);
      }
    }
    if (ok && hi > lo) {
      for (int i : P.slice_ops) {
        PNode& o = P.nodes[P.ops[i].out];
        if (!o.gathered && !o.is_out && o.offset >= 0) o.loop_local = true;
      }
      P.lane_shift = ((P.info.peak_bytes - lo + 255) / 256) * 256;
      P.info.peak_bytes = hi + P.lane_shift;
    }
  }

  // ---- launch schedules + small-op table ------------------------------------------------------------
  auto buf_of = [&](int node, int32_t& buf, int64_t& off) {
    const PNode& nd = P.nodes[node];
    if (nd.is_out) { buf = 1; off = 0; }
    else { buf = 0; off = nd.offset; }
  };
  auto push_small = [&](int i) {
    POp& op = P.ops[i];
    const GenDesc& g = op.gd;
    SmallOp so{};
    so.n_free = g.n_free; so.n_red = g.n_red;
    so.n_out_iter = (int32_t)g.n_out_iter; so.n_red_iter = (int32_t)g.n_red_iter;
    for (int q = 0; q < g.n_free + g.n_red; q++) {
      so.ext[q] = g.ext[q]; so.sl[q] = (int32_t)g.sl[q]; so.sr[q] = (int32_t)g.sr[q]; so.so[q] = (int32_t)g.so[q];
    }
    so.dl = (int8_t)g.dl; so.dr = (int8_t)g.dr; so.dout = (int8_t)g.dout;
    so.clamp = (int8_t)g.clamp; so.has_rhs = (int8_t)g.has_rhs;
    so.binary = 1;
    for (int q = 0; q < g.n_free + g.n_red; q++) if (g.ext[q] != 2) so.binary = 0;
    const PNode& o = P.nodes[op.out];
    so.zero_first = op.gd_zero ? (int32_t)(o.bytes / dtype_size(o.dtype)) : 0;
    buf_of(op.out, so.out_buf, so.out_off);
    if (op.kind == OP_LEAF) {
      so.lhs_buf = 2 + op.leaf; so.lhs_off = 0;
      so.rhs_buf = 0; so.rhs_off = 0;
    } else {
      buf_of(op.lhs, so.lhs_buf, so.lhs_off);
      buf_of(op.rhs, so.rhs_buf, so.rhs_off);
    }
    P.small_ops.push_back(so);
  };
  static const int chain_max = [] { const char* e = getenv("TSB_CHAIN_MAX_OPS"); return e ? atoi(e) : 2; }();  // 0: off
  auto build_schedule = [&](const std::vector<int>& ops, int n_waves, std::vector<Launch>& sched) {
    std::vector<std::vector<int>> small(n_waves), big(n_waves);
    for (int i : ops) (P.ops[i].small ? small : big)[P.ops[i].wave].push_back(i);
    // a wave joins a chain when all it holds is one or two small ops (more are better off one CTA each)
    auto chainable = [&](int w) { return big[w].empty() && !small[w].empty() && (int)small[w].size() <= chain_max; };
    for (int w = 0; w < n_waves; w++) {
      if (chainable(w) && w + 1 < n_waves && chainable(w + 1)) {
        Launch c;
        c.wave = w;
        c.chain = true;
        c.batch_begin = (int)P.small_ops.size();
        while (w < n_waves && chainable(w)) {
          for (int i : small[w]) push_small(i);
          w++;
        }
        w--;
        c.batch_end = (int)P.small_ops.size();
        sched.push_back(c);
        continue;
      }
      Launch batch;
      batch.wave = w;
      batch.batch_begin = (int)P.small_ops.size();
      for (int i : small[w]) push_small(i);
      batch.batch_end = (int)P.small_ops.size();
      {
        // one CTA of 128 threads per op is fine up to a few thousand inner iterations; the larger ops of a wave
        // (up to 2^17) would otherwise take tens of microseconds on the critical path
        int64_t work = 0;
        for (int q = batch.batch_begin; q < batch.batch_end; q++)
          work = std::max<int64_t>(work, (int64_t)P.small_ops[q].n_out_iter * std::max(1, P.small_ops[q].n_red_iter));
        batch.split = (int)std::min<int64_t>(32, std::max<int64_t>(1, work / 2048));
      }
      if (batch.batch_end > batch.batch_begin) sched.push_back(batch);
      for (int i : big[w]) {
        Launch l;
        l.wave = w;
        l.op = i;
        sched.push_back(l);
      }
    }
  };
  build_schedule(pre_ops, n_pre, P.schedule);
  build_schedule(post_ops, n_post, P.post_schedule);
  if (!P.small_ops.empty()) {
    TSB_CUDA(cudaMalloc((void**)&P.d_small_ops, P.small_ops.size() * sizeof(SmallOp)));
    TSB_CUDA(cudaMemcpy(P.d_small_ops, P.small_ops.data(), P.small_ops.size() * sizeof(SmallOp), cudaMemcpyHostToDevice));
  }
  TSB_CUDA(cudaMalloc((void**)&P.d_ptrs, (size_t)(L + 2) * sizeof(void*)));
  P.h_ptrs.assign((size_t)L + 2, nullptr);

  // ---- info ------------------------------------------------------------------------------------
  P.info.n_slices = prod_ext(P.sliced, P.wext);
  P.info.out_numel = prod_ext(P.out_wires, P.wext);
  P.info.out_dtype = P.mode == TSB_SAT ? TSB_U8 : TSB_I64;
  P.info.n_steps = n_steps;
  int mx = 0;
  P.info.n_kernels_shared = (int)(P.schedule.size() + P.post_schedule.size());
  for (auto& op : P.ops) {
    if (op.per_slice) {
      P.info.n_kernels_per_slice++;
      P.info.alg_bytes_per_slice += op.alg_bytes;
      P.info.alg_bytes_u8_per_slice += op.alg_bytes_u8;
      P.info.alg_ops_per_slice += op.alg_ops;
    } else {
      P.info.alg_bytes_shared += op.alg_bytes;
      P.info.alg_bytes_u8_shared += op.alg_bytes_u8;
      P.info.alg_ops_shared += op.alg_ops;
    }
    if (op.kind == OP_STEP) mx = std::max(mx, ilog2(std::max<int64_t>(1, prod_ext(op.ow, P.wext))));
  }
  P.info.log2_max_tensor = mx;
  {
    // per-tile offset tables of the factored fused steps: they depend on the descriptor only, so they are deposited
    // once here instead of by every CTA of every launch (a sixth of the slice loop on the n=50 bench)
    static const bool pre_tab = [] { const char* e = getenv("TSB_BITSF_TABLES"); return !(e && e[0] == '0'); }();
    std::vector<uint32_t> all;
    for (auto& op : P.ops) {
      if (!pre_tab || !op.use_f) continue;
      op.tab_off = (int64_t)all.size();
      const std::vector<uint32_t> t = bits_f_tables(op.fd);
      all.insert(all.end(), t.begin(), t.end());
    }
    if (!all.empty()) {
      TSB_CUDA(cudaMalloc((void**)&P.d_tabs, all.size() * sizeof(uint32_t)));
      TSB_CUDA(cudaMemcpy(P.d_tabs, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
  }
  for (auto& op : P.ops)
    if (op.kind == OP_LEAF && !op.small) P.bakes_leaf_ptrs = true;
  if (P.n_leaves == 1) P.bakes_leaf_ptrs = true;
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

// A rank that executes an aligned block of 2^j slice ids needs, of a pre tensor that keeps sliced wires as leading
// axes, only the sub-block where the fixed (leading) sliced wires take its values: a contiguous, aligned
// 1 / 2^f_bits of the tensor starting at element `off`.  The fixed wires are a prefix of the node's slab.
static void pre_block(const tsb_plan& P, const PNode& nd, int& f_bits, int64_t& off) {
  f_bits = 0;
  off = 0;
  if (P.rr_nfixed == 0 || !nd.shared || nd.post || nd.gathered || nd.is_out) return;
  for (auto w : nd.slab)
    for (int i = 0; i < P.rr_nfixed; i++)
      if (P.sliced[i] == w) f_bits++;
  if (f_bits) off = shared_offset(P, nd, P.rr_digits);
}

void run_step(tsb_plan& P, POp& op, uint8_t* base, void* out_ptr, const std::vector<int64_t>* digits, bool accumulate,
              cudaStream_t stream, int64_t lane_off = 0) {
  tsb_ctx* ctx = P.ctx;
  // address of a node inside the arena block; the second slice lane keeps its own copies of the loop-local nodes
  auto at = [&](const PNode& x) { return base + x.offset + (x.loop_local ? lane_off : 0); };
  // pre ops of a restricted execution: output sub-block (f bits fewer) and the matching operand sub-blocks
  int fo = 0;
  int64_t oo = 0;
  if (!op.per_slice && digits == nullptr) pre_block(P, P.nodes[op.out], fo, oo);
  auto operand_off = [&](const PNode& x) {
    int fx = 0;
    int64_t ox = 0;
    if (fo) pre_block(P, x, fx, ox);
    return ox;
  };
  if (op.kind == OP_PACK || op.kind == OP_UNPACK || op.kind == OP_EXPAND) {
    const PNode& src = P.nodes[op.lhs];
    const PNode& dst = P.nodes[op.out];
    if (op.kind != OP_UNPACK) {
      const int64_t so = operand_off(src);   // entries; src and dst keep the same sliced wires
      if (op.use_bits) {
        MergeDesc d = op.md;
        if (fo && (d.n_vec >> fo) >= (int64_t(1) << d.tile_bits) && (oo & 127) == 0) {
          d.n_vec >>= fo;
          P.rr_launches++;
          launch_merge_bits(ctx, d, at(src), (uint64_t)so, nullptr, 0, at(dst) + (oo >> 3), stream);
        } else {
          launch_merge_bits(ctx, d, at(src), 0, nullptr, 0, at(dst), stream);
        }
      } else if (op.cv_map) {
        const bool src_packed = op.kind == OP_EXPAND;
        const int64_t n = dst.stored_numel();
        if (fo && (n >> fo) >= 1024 && (oo & 127) == 0 && (!src_packed || (so & 31) == 0) && ++P.rr_launches)
          launch_pack_map(ctx, at(src) + (src_packed ? (so >> 3) : so), src_packed, at(dst) + (oo >> 3),
                          n >> fo, op.cv_lo, op.cv_hi, stream);
        else
          launch_pack_map(ctx, at(src), src_packed, at(dst), n, op.cv_lo, op.cv_hi, stream);
      } else {
        const int64_t n = src.stored_numel();
        if (fo && (n >> fo) >= 1024 && (oo & 127) == 0 && ++P.rr_launches)
          launch_pack_bits(ctx, at(src) + oo, at(dst) + (oo >> 3), n >> fo, stream);
        else
          launch_pack_bits(ctx, at(src), at(dst), n, stream);
      }
      return;
    }
    void* pd = dst.is_out ? out_ptr : (void*)(at(dst));
    if (op.per_slice && dst.gathered) pd = (uint8_t*)pd + shared_offset(P, dst, *digits);
    if (fo && (src.stored_numel() >> fo) >= 1024 && (oo & 127) == 0) {
      P.rr_launches++;
      launch_unpack_bits(ctx, at(src) + (oo >> 3), (uint8_t*)pd + oo, src.stored_numel() >> fo, false, stream);
      return;
    }
    launch_unpack_bits(ctx, at(src), pd, src.stored_numel(), accumulate && dst.is_out, stream);
    return;
  }
  const PNode& a = P.nodes[op.lhs];
  const PNode& b = P.nodes[op.rhs];
  const PNode& o = P.nodes[op.out];
  if (op.use_bits && !op.xn.empty()) {
    // fused n-ary step on packed operands
    const void* ptrs[kBitsMaxOps];
    uint64_t offs[kBitsMaxOps];
    const bool sub = fo && op.use_f && (op.fd.n_vec >> fo) >= (int64_t(1) << op.fd.tile_bits) && (oo & 127) == 0;
    for (int j = 0; j < op.n_operands(); j++) {
      const PNode& x = P.nodes[op.operand(j)];
      ptrs[j] = at(x);
      offs[j] = (op.per_slice && x.shared) ? (uint64_t)shared_offset(P, x, *digits) : sub ? (uint64_t)operand_off(x) : 0;
    }
    // a packed gathered result: slice s fills bit block s
    uint8_t* po_bits = at(o) + ((op.per_slice && o.gathered) ? (shared_offset(P, o, *digits) >> 3) : 0);
    if (sub) {
      BitsFDesc d = op.fd;
      d.n_vec >>= fo;
      P.rr_launches++;
      launch_merge_bits_f(ctx, d, ptrs, offs, at(o) + (oo >> 3), stream, op.tab_off >= 0 ? P.d_tabs + op.tab_off : nullptr);
    } else if (op.use_f) {
      launch_merge_bits_f(ctx, op.fd, ptrs, offs, po_bits, stream, op.tab_off >= 0 ? P.d_tabs + op.tab_off : nullptr);
    } else {
      launch_merge_bits_n(ctx, op.nd, ptrs, offs, po_bits, stream);
    }
    return;
  }
  if (op.use_bits) {
    // packed operands: a slice of a shared tensor is addressed by its bit offset
    uint64_t bl = 0, br = 0;
    if (op.per_slice) {
      if (a.shared) bl = (uint64_t)shared_offset(P, a, *digits);
      if (b.shared) br = (uint64_t)shared_offset(P, b, *digits);
    }
    if (fo && (op.md.n_vec >> fo) >= (int64_t(1) << op.md.tile_bits) && (oo & 127) == 0) {
      MergeDesc d = op.md;
      d.n_vec >>= fo;
      P.rr_launches++;
      launch_merge_bits(ctx, d, at(a), (uint64_t)operand_off(a), at(b), (uint64_t)operand_off(b),
                        at(o) + (oo >> 3), stream);
      return;
    }
    launch_merge_bits(ctx, op.md, at(a), bl, at(b), br,
                      at(o) + ((op.per_slice && o.gathered) ? (shared_offset(P, o, *digits) >> 3) : 0), stream);
    return;
  }
  const uint8_t* pl = at(a);
  const uint8_t* pr = at(b);
  if (op.per_slice) {
    if (a.shared) pl += shared_offset(P, a, *digits) * dtype_size(a.dtype);
    if (b.shared) pr += shared_offset(P, b, *digits) * dtype_size(b.dtype);
  }
  void* po = o.is_out ? out_ptr : (void*)(at(o));
  if (op.per_slice && o.gathered) po = (uint8_t*)po + shared_offset(P, o, *digits) * dtype_size(o.dtype);
  bool acc = accumulate && o.is_out;
  if (P.stack_out && op.out == P.root) {
    // root result goes into the slice's block of the stacked result buffer, accumulating
    const PNode& r = P.nodes[P.rnode];
    po = at(r) + shared_offset(P, r, *digits) * dtype_size(r.dtype);
    acc = true;
  }
  const bool clamp = P.mode == TSB_SAT;
  int dout = o.dtype;
  if (P.stack_out && op.out == P.root) dout = P.nodes[P.rnode].dtype;
  // each operand only has to be aligned to the vector it is loaded with (slices of tiny shared tensors
  // start at multiples of their own size), the result to 16 bytes
  const uintptr_t al = ((uintptr_t)dtype_size(a.dtype) << __builtin_popcount((unsigned)op.md.nibL)) - 1;
  const uintptr_t ar = ((uintptr_t)dtype_size(b.dtype) << __builtin_popcount((unsigned)op.md.nibR)) - 1;
  if (op.use_merge && ((uintptr_t)pl & al) == 0 && ((uintptr_t)pr & ar) == 0 && ((uintptr_t)po & 15) == 0) {
    MergeDesc d = op.md;
    d.accumulate = acc ? 1 : 0;
    if (fo && !o.is_out && !acc && (d.n_vec >> fo) >= (int64_t(1) << d.tile_bits)) {
      const uint8_t* ql = pl + operand_off(a) * dtype_size(a.dtype);
      const uint8_t* qr = pr + operand_off(b) * dtype_size(b.dtype);
      uint8_t* qo = (uint8_t*)po + oo * dtype_size(dout);
      if (((uintptr_t)ql & al) == 0 && ((uintptr_t)qr & ar) == 0 && ((uintptr_t)qo & 15) == 0) {
        d.n_vec >>= fo;
        P.rr_launches++;
        launch_merge(ctx, d, ql, a.dtype, qr, b.dtype, qo, dout, clamp, stream);
        return;
      }
    }
    launch_merge(ctx, d, pl, a.dtype, pr, b.dtype, po, dout, clamp, stream);
  } else if (op.has_gd && !acc && dout == o.dtype) {
    if (op.gd_zero) TSB_CUDA(cudaMemsetAsync(po, 0, o.bytes, stream));
    launch_generic(ctx, op.gd, pl, pr, po, stream);
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
  if (op.has_gd) {
    if (op.gd_zero) TSB_CUDA(cudaMemsetAsync(po, 0, o.bytes, stream));
    launch_generic(P.ctx, op.gd, leaf_ptr, nullptr, po, stream);
    return;
  }
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
    ctx->refs++;   // a plan keeps its context alive (it returns arena blocks to it when destroyed)
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

// operand fetches (16 bytes each) per four output vectors of a fused step
static int fetches_x4(const POp& op) {
  if (op.xn.empty()) return 0;
  if (!op.use_f) return 4 * op.nd.n_ops << op.nd.kbits;
  int n = 0;
  for (int s = 0; s < op.fd.n_ops; s++) {
    const int kk = 1 << __builtin_popcount((unsigned)op.fd.op[s].kmask);
    if (op.fd.n_pair >= 2) {   // private streaming: slot 0 once per 1 / 2 / 4 vectors, slot 1 per 4 / 8, the others per all of an item's (rounded up)
      const int sh = s == 0 ? op.fd.n_pair2 : s == 1 ? op.fd.n_pair : op.fd.n_pair + op.fd.n_pair2;
      n += (4 * kk + (1 << sh) - 1) >> sh;
    } else {
      n += ((4 >> op.fd.n_pair) * kk) << __builtin_popcount((unsigned)op.fd.op[s].held);
    }
  }
  return n;
}

int tsb_plan_describe(const tsb_plan* plan, char* buf, size_t cap, size_t* needed) {
  if (!plan || !needed) { tsb::set_error("null argument"); return TSB_ERR_INVALID; }
  std::ostringstream os;
  os << "{";
  if (plan->phase_valid && plan->phase_ev[3] && cudaEventQuery(plan->phase_ev[3]) == cudaSuccess) {
    // device times of the last execution with phase timing: [once-only schedule | slice loop | tail + combine]
    float pre = 0, loop = 0, post = 0;
    cudaEventElapsedTime(&pre, plan->phase_ev[0], plan->phase_ev[1]);
    cudaEventElapsedTime(&loop, plan->phase_ev[1], plan->phase_ev[2]);
    cudaEventElapsedTime(&post, plan->phase_ev[2], plan->phase_ev[3]);
    os << "\"phase_ms\":{\"pre\":" << pre << ",\"slice_loop\":" << loop << ",\"post\":" << post << "},";
  }
  os << "\"peak_bytes\":" << plan->info.peak_bytes << ",\"slice_lanes\":" << (plan->lane_shift > 0 ? 2 : 1) << ",\"n_slices\":" << plan->info.n_slices
     << ",\"restricted_launches\":" << plan->rr_launches << ",\"graph_captures\":" << plan->n_captures
     << ",\"graph_replays\":" << plan->n_replays << ",\"ops\":[";
  bool first = true;
  for (auto& op : plan->ops) {
    if (!first) os << ",";
    first = false;
    const PNode& o = plan->nodes[op.out];
    os << "{\"kind\":\"" << (op.kind == OP_LEAF ? "leaf" : op.kind == OP_PACK ? "pack" : op.kind == OP_UNPACK ? "unpack" : op.kind == OP_EXPAND ? "expand" : op.kind == OP_STEP ? (op.use_bits ? "bits" : op.use_gemm ? "gemm" : op.use_merge ? "merge" : "generic") : "permute")
       << "\",\"per_slice\":" << (op.per_slice ? 1 : 0) << ",\"post\":" << (plan->nodes[op.out].post ? 1 : 0) << ",\"out_bits\":" << ilog2(std::max<int64_t>(1, o.reg_numel))
       << ",\"dtype\":" << o.dtype << ",\"b\":" << op.b << ",\"m\":" << op.m << ",\"n\":" << op.n << ",\"k\":" << op.k
       << ",\"wave\":" << op.wave << ",\"small\":" << (op.small ? 1 : 0) << ",\"factored\":" << (op.use_f ? 1 : 0) << ",\"n_pair\":" << (op.use_f ? op.fd.n_pair : 0) << ",\"n_pair2\":" << (op.use_f ? op.fd.n_pair2 : 0) << ",\"fetches_x4\":" << fetches_x4(op) << ",\"wf\":" << op.wf_chosen << ",\"wf_default\":" << op.wf_default << ",\"out_gathered\":" << (o.gathered ? 1 : 0) << ",\"out_packed\":" << (o.packed ? 1 : 0) << ",\"fused\":" << op.n_fused << ",\"n_operands\":" << op.n_operands() << ",\"nibL\":" << op.md.nibL << ",\"nibR\":" << op.md.nibR
       << ",\"bytes\":" << op.alg_bytes << ",\"ops\":" << op.alg_ops << ",\"ms\":" << op.last_ms;
    if (op.kind == OP_STEP && op.ow.size() >= 20) {
      auto arr = [&](const char* name, const std::vector<int32_t>& v) {
        os << ",\"" << name << "\":[";
        for (size_t q = 0; q < v.size(); q++) os << (q ? "," : "") << v[q];
        os << "]";
      };
      arr("lw", op.lw); arr("rw", op.rw); arr("ow", op.ow);
      for (size_t x = 0; x < op.xw.size(); x++) arr(x == 0 ? "xw0" : x == 1 ? "xw1" : "xw2", op.xw[x]);
    }
    os << "}";
  }
  os << "],\"launches\":[";
  first = true;
  std::vector<Launch> all = plan->schedule;
  all.insert(all.end(), plan->post_schedule.begin(), plan->post_schedule.end());
  for (auto& l : all) {
    if (!first) os << ",";
    first = false;
    os << "{\"wave\":" << l.wave << ",\"op\":" << l.op << ",\"n_small\":" << (l.batch_end - l.batch_begin) << ",\"chain\":" << (l.chain ? 1 : 0)
       << ",\"ms\":" << l.last_ms << "}";
  }
  os << "]}";
  const std::string s = os.str();
  *needed = s.size() + 1;
  if (buf && cap >= s.size() + 1) memcpy(buf, s.c_str(), s.size() + 1);
  return TSB_OK;
}

int tsb_plan_set_timing(tsb_plan* plan, int enabled) {
  if (!plan) { tsb::set_error("null argument"); return TSB_ERR_INVALID; }
  plan->timing = enabled < 0 ? 0 : enabled > 2 ? 1 : enabled;
  return TSB_OK;
}

int tsb_plan_timing_json(tsb_plan* plan, char* buf, size_t cap, size_t* needed) {
  return tsb_plan_describe(plan, buf, cap, needed);
}

}  // extern "C"

namespace {

// the slice mask of a masked root as a range of ones, when it is one (it is whenever the root's leading axes are
// all the sliced wires in slice-id order, i.e. always unless a sliced wire touches no tensor)
struct MaskPlan {
  std::vector<uint8_t> host;   // full mask (fallback: uploaded from the host, direct executions only)
  int64_t lo = 0, hi = 0;      // ones at [lo, hi)
  bool is_range = false;
};

MaskPlan make_mask(const tsb_plan& P, int64_t slice_begin, int64_t slice_end) {
  MaskPlan m;
  if (!P.mask_root) return m;
  const PNode& r = P.nodes[P.root];
  m.host.assign((size_t)r.s_numel, 0);
  for (int64_t sid = slice_begin; sid < slice_end; sid++) {
    const std::vector<int64_t> digits = slice_digits(P, sid);
    m.host[(size_t)(shared_offset(P, r, digits) / r.reg_numel)] = 1;
  }
  int64_t lo = 0, hi = (int64_t)m.host.size();
  while (lo < hi && !m.host[(size_t)lo]) lo++;
  while (hi > lo && !m.host[(size_t)hi - 1]) hi--;
  m.is_range = true;
  for (int64_t i = lo; i < hi; i++) m.is_range = m.is_range && m.host[(size_t)i];
  m.lo = lo; m.hi = hi;
  return m;
}

// Enqueues one whole execution on `stream`: [once-only schedule][slice loop][tail schedule][combine].  Everything it
// launches depends only on (plan, base, out, slice range) — and on the leaf pointers where P.bakes_leaf_ptrs — so
// the same call can run directly or under stream capture.  The pointer table d_ptrs is uploaded by the caller.
void enqueue_execution(tsb_plan& P, uint8_t* base, const void* const* leaf_ptrs, int64_t slice_begin, int64_t slice_end,
                       void* out, const MaskPlan& mask, cudaStream_t stream, int timing_mode, int n_side_max) {
  tsb_ctx* ctx = P.ctx;
  const bool timing = timing_mode == 1;
  const bool phases = timing_mode == 2;
  auto phase_mark = [&](int i) {
    if (!phases) return;
    if (!P.phase_ev[i]) TSB_CUDA(cudaEventCreate(&P.phase_ev[i]));
    TSB_CUDA(cudaEventRecord(P.phase_ev[i], stream));
  };
  P.phase_valid = false;
  phase_mark(0);
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
  for (int g : P.gathered_nodes) TSB_CUDA(cudaMemsetAsync(base + P.nodes[g].offset, 0, P.nodes[g].bytes, stream));
  // An op may run on a side stream when its launch needs nothing but its descriptor and the arena block: packed
  // steps, conversions and aligned byte merges (the other paths take temporaries from the stream-ordered arena).
  auto forkable = [&](const Launch& l) {
    if (l.op < 0) return false;
    const POp& op = P.ops[l.op];
    if (op.kind == OP_LEAF || P.nodes[op.out].is_out) return false;
    if (op.kind == OP_PACK || op.kind == OP_UNPACK || op.kind == OP_EXPAND) return true;
    if (op.kind != OP_STEP) return false;
    if (op.use_bits) return true;
    if (!op.use_merge || op.use_gemm) return false;
    const PNode& a = P.nodes[op.lhs];
    const PNode& b = P.nodes[op.rhs];
    const PNode& o = P.nodes[op.out];
    const uintptr_t al = ((uintptr_t)dtype_size(a.dtype) << __builtin_popcount((unsigned)op.md.nibL)) - 1;
    const uintptr_t ar = ((uintptr_t)dtype_size(b.dtype) << __builtin_popcount((unsigned)op.md.nibR)) - 1;
    return !a.is_out && !b.is_out && ((uintptr_t)a.offset & al) == 0 && ((uintptr_t)b.offset & ar) == 0 &&
           ((uintptr_t)o.offset & 15) == 0 && P.rr_nfixed == 0;
  };
  auto run_one = [&](Launch& l, cudaStream_t st) {
    l.last_ms = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (timing) {
      TSB_CUDA(cudaEventCreate(&t0));
      TSB_CUDA(cudaEventCreate(&t1));
      TSB_CUDA(cudaEventRecord(t0, st));
    }
    if (l.op < 0) {
      const int n = l.batch_end - l.batch_begin;
      if (l.chain)
        TSB_LAUNCH_PDL(k_small_chain, dim3(1), dim3(kChainThreads), 0, st,
                       (const SmallOp*)(P.d_small_ops + l.batch_begin), n, (void* const*)P.d_ptrs);
      else
        TSB_LAUNCH_PDL(k_small_batch, dim3(n, l.split), dim3(kSmallThreads), 0, st,
                       (const SmallOp*)(P.d_small_ops + l.batch_begin), (void* const*)P.d_ptrs);
      ctx->note(l.chain ? "k_small_chain" : "k_small_batch");
      TSB_CUDA(cudaGetLastError());
    } else {
      POp& op = P.ops[l.op];
      if (op.kind == OP_LEAF) {
        if (!(P.n_leaves == 1 && !P.sliced.empty())) run_leaf(P, op, base, leaf_ptrs[op.leaf], out, st);
      } else {
        run_step(P, op, base, out, nullptr, false, st);
      }
    }
    if (timing) {
      TSB_CUDA(cudaEventRecord(t1, st));
      TSB_CUDA(cudaEventSynchronize(t1));
      float ms = 0;
      TSB_CUDA(cudaEventElapsedTime(&ms, t0, t1));
      l.last_ms = ms;
      if (l.op >= 0) P.ops[l.op].last_ms = ms;
      cudaEventDestroy(t0);
      cudaEventDestroy(t1);
    }
  };
  // once-only parts: waves of independent ops (small ops of a wave = one launch; the other launches of a wave are
  // spread over the caller's stream and a few side streams, joined before the next wave — under capture this
  // becomes parallel branches of the graph at no host cost per replay)
  auto run_schedule = [&](std::vector<Launch>& sched) {
    size_t li = 0;
    while (li < sched.size()) {
      size_t lj = li;
      if (!timing && n_side_max > 0 && forkable(sched[li]))
        while (lj < sched.size() && sched[lj].wave == sched[li].wave && forkable(sched[lj])) lj++;
      const int group = (int)(lj - li);
      if (group < 3) {
        run_one(sched[li], stream);
        li++;
        continue;
      }
      const int n_side = std::min(n_side_max, group - 1);
      while ((int)P.side.size() < n_side) {
        cudaStream_t st;
        TSB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        P.side.push_back(st);
      }
      while ((int)P.side_ev.size() < 1 + n_side) {
        cudaEvent_t e;
        TSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        P.side_ev.push_back(e);
      }
      TSB_CUDA(cudaEventRecord(P.side_ev[0], stream));
      for (int q = 0; q < n_side; q++) TSB_CUDA(cudaStreamWaitEvent(P.side[q], P.side_ev[0], 0));
      for (int g = 0; g < group; g++) {
        const int lane = g % (n_side + 1);
        run_one(sched[li + g], lane == 0 ? stream : P.side[lane - 1]);
      }
      for (int q = 0; q < n_side; q++) {
        TSB_CUDA(cudaEventRecord(P.side_ev[1 + q], P.side[q]));
        TSB_CUDA(cudaStreamWaitEvent(stream, P.side_ev[1 + q], 0));
      }
      li = lj;
    }
  };
  // A rank that executes an aligned power-of-two block of slice ids computes only its own sub-blocks of the
  // once-only tensors that keep sliced wires (run_step / pre_block); everything else is replicated.
  P.rr_nfixed = 0;
  P.rr_launches = 0;
  {
    static const bool restrict_pre = [] { const char* e = getenv("TSB_PRE_RESTRICT"); return !(e && e[0] == '0'); }();
    const int64_t n = slice_end - slice_begin, total = P.info.n_slices;
    bool binary = !P.sliced.empty();
    for (auto w : P.sliced) binary = binary && P.wext[w] == 2;
    if (restrict_pre && binary && n > 0 && n < total && (n & (n - 1)) == 0 && slice_begin % n == 0) {
      P.rr_nfixed = (int)P.sliced.size() - ilog2(n);
      P.rr_digits = slice_digits(P, slice_begin);
    }
  }
  run_schedule(P.schedule);
  P.rr_nfixed = 0;
  phase_mark(1);
  // per-slice part
  if (!P.sliced.empty()) {
    const bool lanes = P.lane_shift > 0 && !timing && n_side_max > 0 && slice_end - slice_begin >= 2;
    if (lanes) {
      if (!P.lane_stream) TSB_CUDA(cudaStreamCreateWithFlags(&P.lane_stream, cudaStreamNonBlocking));
      for (auto& e : P.lane_ev)
        if (!e) TSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      TSB_CUDA(cudaEventRecord(P.lane_ev[0], stream));
      TSB_CUDA(cudaStreamWaitEvent(P.lane_stream, P.lane_ev[0], 0));
    }
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
      // two lanes: slices alternate between the caller's stream and the plan's lane stream (captured: two branches)
      const bool second = lanes && ((sid - slice_begin) & 1);
      for (int i : P.slice_ops) {
        POp& op = P.ops[i];
        timed((size_t)i, [&] { run_step(P, op, base, out, &digits, sid != slice_begin, second ? P.lane_stream : stream, second ? P.lane_shift : 0); });
      }
    }
    if (lanes) {
      TSB_CUDA(cudaEventRecord(P.lane_ev[1], P.lane_stream));
      TSB_CUDA(cudaStreamWaitEvent(stream, P.lane_ev[1], 0));
    }
  }
  phase_mark(2);
  run_schedule(P.post_schedule);
  if (P.mask_root) {
    // out[o] = sum over sliced-wire combinations s in [slice_begin, slice_end) of root[s, o]
    const PNode& r = P.nodes[P.root];
    uint8_t* d_mask = base + P.mask_offset;
    if (mask.is_range) {
      TSB_CUDA(cudaMemsetAsync(d_mask, 0, mask.host.size(), stream));
      if (mask.hi > mask.lo) TSB_CUDA(cudaMemsetAsync(d_mask + mask.lo, 1, (size_t)(mask.hi - mask.lo), stream));
    } else {
      TSB_CUDA(cudaMemcpyAsync(d_mask, mask.host.data(), mask.host.size(), cudaMemcpyHostToDevice, stream));
    }
    const std::vector<int32_t> rw = r.stored();
    contract2_impl(ctx, base + r.offset, (int)rw.size(), rw.data(), r.dtype, d_mask, (int)r.slab.size(), r.slab.data(),
                   TSB_U8, out, (int)r.reg.size(), r.reg.data(), P.info.out_dtype, P.n_wires, P.wext.data(), P.mode, false,
                   false, stream);
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
  phase_mark(3);
  P.phase_valid = phases;
}

// pointer table for the batched small ops: [arena, out, leaves...]; uploaded only when it changed
void upload_ptrs(tsb_plan& P, uint8_t* base, void* out, const void* const* leaf_ptrs, cudaStream_t stream) {
  if (P.small_ops.empty()) return;
  P.h_ptrs[0] = base;
  P.h_ptrs[1] = out;
  for (int i = 0; i < P.n_leaves; i++) P.h_ptrs[2 + i] = const_cast<void*>(leaf_ptrs[i]);
  if (P.h_ptrs_uploaded == P.h_ptrs) return;
  // pageable source: the runtime stages it before returning, so h_ptrs may change right after
  TSB_CUDA(cudaMemcpyAsync(P.d_ptrs, P.h_ptrs.data(), (size_t)(P.n_leaves + 2) * sizeof(void*), cudaMemcpyHostToDevice,
                           stream));
  P.h_ptrs_uploaded = P.h_ptrs;
}

#ifndef TSB_HOST_EMU
void destroy_graph_entry(tsb_plan& P, tsb_plan::GraphEntry& e) {
  if (e.exec) cudaGraphExecDestroy(e.exec);
  for (void* p : e.held.owned) P.ctx->arena.release(p);
  e.held.owned.clear();
  e.held.free.clear();
  e.exec = nullptr;
}

constexpr size_t kGraphSmallResult = 64 * 1024;  // results up to this size go through the plan's own buffer
constexpr size_t kGraphCacheEntries = 8;

// Replays (capturing first, on the second execution with the same key) the execution as a CUDA graph.
// Returns false when this call must run directly.
bool execute_as_graph(tsb_plan& P, const void* const* leaf_ptrs, int64_t slice_begin, int64_t slice_end, void* out,
                      const MaskPlan& mask, cudaStream_t stream) {
  static const bool enabled = [] { const char* e = getenv("TSB_GRAPH"); return !(e && e[0] == '0'); }();
  static const int graph_streams = [] { const char* e = getenv("TSB_GRAPH_STREAMS"); return e ? atoi(e) : 4; }();
  tsb_ctx* ctx = P.ctx;
  if (!enabled || (P.flags & TSB_PLAN_NO_GRAPH) || P.timing) return false;
  if (P.mask_root && !mask.is_range) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
  const size_t out_bytes = (size_t)P.info.out_numel * dtype_size(P.info.out_dtype);
  const bool own_result = out_bytes <= kGraphSmallResult;
  tsb_plan::GraphKey key{own_result ? nullptr : out, slice_begin, slice_end, 0};
  if (P.bakes_leaf_ptrs) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < P.n_leaves; i++) h = (h ^ (uint64_t)(uintptr_t)leaf_ptrs[i]) * 1099511628211ull;
    key.leaf_hash = h;
  }
  if (P.ungraphable.count(key)) return false;
  auto it = P.graphs.find(key);
  if (it == P.graphs.end()) {
    if (P.seen[key]++ < 1) return false;  // first execution: direct (also warms the arena and the module)
    // ---- capture
    if (!P.gblock) P.gblock = ctx->arena.alloc((size_t)std::max<int64_t>(P.info.peak_bytes, 256));
    if (own_result && !P.gres) P.gres = ctx->arena.alloc(std::max<size_t>(out_bytes, 256));
    if (P.graphs.size() >= kGraphCacheEntries) {  // evict the least recently used entry
      auto victim = P.graphs.begin();
      for (auto q = P.graphs.begin(); q != P.graphs.end(); ++q)
        if (q->second.last_used < victim->second.last_used) victim = q;
      if (P.has_last) cudaStreamSynchronize(P.last_stream);
      destroy_graph_entry(P, victim->second);
      P.graphs.erase(victim);
    }
    tsb_plan::GraphEntry e;
    const int64_t launches0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    if (!P.cap_stream && cudaStreamCreateWithFlags(&P.cap_stream, cudaStreamNonBlocking) != cudaSuccess) P.cap_stream = nullptr;
    bool ok = P.cap_stream && cudaStreamBeginCapture(P.cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      capture_state().hold = &e.held;
      capture_state().counts = &e.kcount;
      try {
        enqueue_execution(P, (uint8_t*)P.gblock, leaf_ptrs, slice_begin, slice_end, own_result ? P.gres : out, mask,
                          P.cap_stream, false, graph_streams);
      } catch (const tsb::Error&) {
        ok = false;
      }
      capture_state() = CaptureState{};
      if (cudaStreamEndCapture(P.cap_stream, &graph) != cudaSuccess || graph == nullptr) ok = false;
    }
    if (ok && cudaGraphInstantiate(&e.exec, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    // the captured launches did not run: they are counted per replay
    e.launches = ctx->launches - launches0;
    ctx->launches = launches0;
    if (ctx->trace)
      for (auto& kv : e.kcount) ctx->kcount[kv.first] -= kv.second;
    e.rr_launches = P.rr_launches;
    if (!ok) {
      (void)cudaGetLastError();
      destroy_graph_entry(P, e);
      P.ungraphable.insert(key);
      return false;
    }
    P.n_captures++;
    it = P.graphs.emplace(key, std::move(e)).first;
  }
  tsb_plan::GraphEntry& e = it->second;
  e.last_used = ++P.tick;
  upload_ptrs(P, (uint8_t*)P.gblock, own_result ? P.gres : out, leaf_ptrs, stream);
  TSB_CUDA(cudaGraphLaunch(e.exec, stream));
  if (own_result) TSB_CUDA(cudaMemcpyAsync(out, P.gres, out_bytes, cudaMemcpyDeviceToDevice, stream));
  ctx->launches += e.launches;
  if (ctx->trace)
    for (auto& kv : e.kcount) ctx->kcount[kv.first] += kv.second;
  P.rr_launches = e.rr_launches;
  P.n_replays++;
  return true;
}
#endif

}  // namespace

extern "C" {

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
    // one execution at a time (the pointer table, the slice-range restriction and the graphs' arena block are
    // per-plan state); an execution on another stream than the previous one is ordered after it
    std::lock_guard<std::mutex> lock(P.mu);
    if (!P.done_ev) TSB_CUDA(cudaEventCreateWithFlags(&P.done_ev, cudaEventDisableTiming));
    if (P.has_last && P.last_stream != stream) TSB_CUDA(cudaStreamWaitEvent(stream, P.done_ev, 0));
    auto mark_done = [&] {
      TSB_CUDA(cudaEventRecord(P.done_ev, stream));
      P.last_stream = stream;
      P.has_last = true;
    };
    const MaskPlan mask = make_mask(P, slice_begin, slice_end);
#ifndef TSB_HOST_EMU
    if (execute_as_graph(P, leaf_ptrs, slice_begin, slice_end, out, mask, stream)) {
      mark_done();
      return TSB_OK;
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool outer_capture = cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    if (outer_capture) {
      // the caller is capturing this stream into its own graph: the launches must keep their memory for as long
      // as the plan lives (the caller's replays use the same addresses), and no event may be recorded
      if (!P.gblock) P.gblock = ctx->arena.alloc((size_t)std::max<int64_t>(P.info.peak_bytes, 256));
      static tsb::HeldBlocks forever;      // temporaries of caller-captured launches stay theirs (recycled in capture order)
      static std::mutex forever_mu;
      std::lock_guard<std::mutex> forever_lock(forever_mu);
      capture_state().hold = &forever;
      try {
        upload_ptrs(P, (uint8_t*)P.gblock, out, leaf_ptrs, stream);
        enqueue_execution(P, (uint8_t*)P.gblock, leaf_ptrs, slice_begin, slice_end, out, mask, stream, false, 0);
      } catch (...) {
        capture_state().hold = nullptr;
        throw;
      }
      capture_state().hold = nullptr;
      return TSB_OK;
    }
#endif
    static const int n_side_max = [] { const char* e = getenv("TSB_WAVE_STREAMS"); return e ? atoi(e) : 0; }();
    Temp block(ctx, (size_t)std::max<int64_t>(P.info.peak_bytes, 256), stream);
    upload_ptrs(P, (uint8_t*)block.p, out, leaf_ptrs, stream);
    enqueue_execution(P, (uint8_t*)block.p, leaf_ptrs, slice_begin, slice_end, out, mask, stream, P.timing, n_side_max);
    mark_done();
    return TSB_OK;
  } catch (const tsb::Error& e) {
    return e.code;
  } catch (const std::exception& e) {
    tsb::set_error("internal error: %s", e.what());
    return TSB_ERR_INVALID;
  }
}

int tsb_plan_execute_host(tsb_plan* plan, const void* host_leaves, size_t leaves_bytes, int64_t slice_begin,
                          int64_t slice_end, void* host_out, size_t out_bytes, tsb_stream stream_) {
  try {
    TSB_CHECK(plan && host_leaves && host_out, TSB_ERR_INVALID, "null argument");
    tsb_plan& P = *plan;
    cudaStream_t stream = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> host_lock(P.hx_mu);
    {
      if (P.hx_ptrs.empty()) {
        // staging buffers of the host entry point, once per plan: leaves back to back, result
        size_t total = 0;
        std::vector<size_t> offs;
        for (const auto& lw : P.leaf_wires) {
          offs.push_back(total);
          total += (size_t)prod_ext(lw, P.wext);
        }
        P.hx_total = total;
        P.hx_leaves = (uint8_t*)P.ctx->arena.alloc(std::max<size_t>(total, 16));
        P.hx_out = P.ctx->arena.alloc(std::max<size_t>((size_t)P.info.out_numel * dtype_size(P.info.out_dtype), 16));
        for (size_t o : offs) P.hx_ptrs.push_back(P.hx_leaves + o);
      }
    }
    const size_t want_out = (size_t)P.info.out_numel * dtype_size(P.info.out_dtype);
    TSB_CHECK(leaves_bytes == P.hx_total, TSB_ERR_INVALID, "host leaves hold %zu bytes, the plan's leaves %zu", leaves_bytes, P.hx_total);
    TSB_CHECK(out_bytes == want_out, TSB_ERR_INVALID, "host result buffer holds %zu bytes, the plan's result %zu", out_bytes, want_out);
    TSB_CUDA(cudaMemcpyAsync(P.hx_leaves, host_leaves, leaves_bytes, cudaMemcpyHostToDevice, stream));
    const int rc = tsb_plan_execute(plan, P.hx_ptrs.data(), slice_begin, slice_end, P.hx_out, stream_);
    if (rc != TSB_OK) return rc;
    TSB_CUDA(cudaMemcpyAsync(host_out, P.hx_out, out_bytes, cudaMemcpyDeviceToHost, stream));
    TSB_CUDA(cudaStreamSynchronize(stream));
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
#ifndef TSB_HOST_EMU
    if (plan->has_last && !plan->graphs.empty()) cudaStreamSynchronize(plan->last_stream);
    for (auto& kv : plan->graphs) destroy_graph_entry(*plan, kv.second);
#endif
    try {
      if (plan->gblock) plan->ctx->arena.release(plan->gblock, plan->has_last ? plan->last_stream : tsb::Arena::no_stream());
      if (plan->gres) plan->ctx->arena.release(plan->gres, plan->has_last ? plan->last_stream : tsb::Arena::no_stream());
      if (plan->hx_leaves) plan->ctx->arena.release(plan->hx_leaves, plan->has_last ? plan->last_stream : tsb::Arena::no_stream());
      if (plan->hx_out) plan->ctx->arena.release(plan->hx_out, plan->has_last ? plan->last_stream : tsb::Arena::no_stream());
    } catch (...) {
    }
    if (plan->done_ev) cudaEventDestroy(plan->done_ev);
    if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
    if (plan->lane_stream) cudaStreamDestroy(plan->lane_stream);
    for (auto e : plan->lane_ev) if (e) cudaEventDestroy(e);
    for (auto e : plan->ev) cudaEventDestroy(e);
    for (auto e : plan->phase_ev) if (e) cudaEventDestroy(e);
    for (auto e : plan->side_ev) cudaEventDestroy(e);
    for (auto st : plan->side) cudaStreamDestroy(st);
    if (plan->d_small_ops) cudaFree(plan->d_small_ops);
    if (plan->d_tabs) cudaFree(plan->d_tabs);
    if (plan->d_ptrs) cudaFree(plan->d_ptrs);
    tsb::ctx_unref(plan->ctx);
    delete plan;
  }
  return TSB_OK;
}

}  // extern "C"
