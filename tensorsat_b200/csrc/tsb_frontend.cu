// tsb_frontend.cu — DIMACS / clause lists straight to the flat relation network, its contraction tree and a plan
// description, without Python diagram objects.  Host code only.
//
// Reference behaviour restated here (paths relative to the reference repository):
//   * DIMACS dialect            tensorsat/lib/sat.py:80-100   ("c" comment lines, "p cnf <vars> <clauses>", one clause
//                                                              per line, single-space separated, trailing 0 dropped)
//   * bintree clause gadget     tensorsat/lib/sat.py:248-262  (NOT gates for negative literals in literal order, a
//                                                              balanced tree of OR gates, bit_1 on its root)
//   * input boxes               tensorsat/lib/sat.py:165-206  (bit_0 / bit_1 / bit_unk, one per variable)
//   * wire and slot numbering of `(cnf.inputs(x) >> cnf.diagram()).flatten()`:
//                               tensorsat/diagrams.py:387-461,1452-1481 (wires 0..n-1 are the variables, every gate
//                               output takes the next wire id in creation order; slots: the input boxes, then per
//                               clause its NOT gates, OR gates and the closing bit_1)
// The numbering is pinned by the wiring digests of tests/golden/instances.json (recorded from the reference).
//
// The tree finder is the native twin of tensorsat_b200/planner/search.py (randomised min-fill / min-degree
// elimination orders, bucket elimination, slicing by weight above the target): the stand-in for the tree cotengra
// hands the reference (tensorsat/contractions/cotengra.py:78-85).  Trials 0 and 1 are deterministic and produce the
// same SSA path as the Python planner (tested); the randomised trials use their own generator.
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "tsb_common.h"

struct tsb_network {
  int32_t n_vars = 0;
  std::vector<std::vector<int32_t>> clauses;
  bool has_inputs = false;
  std::string inputs;                              // '0' / '1' / '?' per variable
  int32_t n_wires = 0;
  std::vector<std::vector<int32_t>> leaf_wires;    // slot order
  std::vector<int32_t> leaf_kind;                  // tsb_leaf_kind
  std::vector<int32_t> out_wires;                  // open variable wires when there are no input boxes
};

struct tsb_tree {
  int32_t n_leaves = 0;
  std::vector<int32_t> ssa;      // 2 * (n_leaves - 1)
  std::vector<int32_t> sliced;
  double log2_width = 0, log2_ops = 0, log2_write = 0, log2_width_unsliced = 0;
};

namespace tsb {
namespace {

#define TSB_FE_BEGIN try {
#define TSB_FE_END                                           \
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

void build_wiring(tsb_network& N) {
  N.leaf_wires.clear();
  N.leaf_kind.clear();
  N.out_wires.clear();
  int32_t next = N.n_vars;
  if (N.has_inputs) {
    for (int v = 0; v < N.n_vars; v++) {
      const char c = N.inputs[(size_t)v];
      N.leaf_wires.push_back({v});
      N.leaf_kind.push_back(c == '0' ? TSB_LEAF_BIT_0 : c == '1' ? TSB_LEAF_BIT_1 : TSB_LEAF_BIT_UNK);
    }
  } else {
    for (int v = 0; v < N.n_vars; v++) N.out_wires.push_back(v);
  }
  for (const auto& clause : N.clauses) {
    std::vector<int32_t> level;
    for (int32_t lit : clause) {
      if (lit > 0) {
        level.push_back(lit - 1);
      } else {
        N.leaf_wires.push_back({-lit - 1, next});
        N.leaf_kind.push_back(TSB_LEAF_NOT);
        level.push_back(next++);
      }
    }
    TSB_CHECK(!level.empty(), TSB_ERR_INVALID, "empty clause");   // (the reference's builder fails on level[0] too)
    while (level.size() > 1) {
      const size_t pairs = level.size() / 2;
      std::vector<int32_t> merged;
      for (size_t p = 0; p < pairs; p++) {
        N.leaf_wires.push_back({level[2 * p], level[2 * p + 1], next});
        N.leaf_kind.push_back(TSB_LEAF_OR);
        merged.push_back(next++);
      }
      if (level.size() % 2) {
        N.leaf_wires.push_back({merged.back(), level.back(), next});
        N.leaf_kind.push_back(TSB_LEAF_OR);
        merged.back() = next++;
      }
      level.swap(merged);
    }
    N.leaf_wires.push_back({level[0]});
    N.leaf_kind.push_back(TSB_LEAF_BIT_1);
  }
  N.n_wires = next;
}

void check_inputs(const tsb_network& N, const char* inputs) {
  TSB_CHECK(std::strlen(inputs) == (size_t)N.n_vars, TSB_ERR_INVALID,
            "Expected input specification string to be %d characters long, found %zu characters instead.", N.n_vars,
            std::strlen(inputs));
  for (const char* p = inputs; *p; p++)
    TSB_CHECK(*p == '0' || *p == '1' || *p == '?', TSB_ERR_INVALID, "Input specification characters must be '0', '1' or '?'.");
}

// ---- splitmix64: the generator of the randomised trials
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// Greedy vertex elimination on the wire primal graph (planner/search.py::_elimination_order).  Adjacency is kept as
// one bit row per wire: the fill-in of a wire is a handful of AND + popcount passes over its neighbours' rows.
std::vector<int32_t> elimination_order(const tsb_network& N, bool min_fill, double noise, Rng& rng) {
  const int W = N.n_wires;
  const int NW = (W + 63) / 64;
  std::vector<uint64_t> adj((size_t)W * NW, 0);
  auto row = [&](int32_t w) { return adj.data() + (size_t)w * NW; };
  auto set_bit = [&](int32_t a, int32_t b) { row(a)[b >> 6] |= uint64_t(1) << (b & 63); };
  std::vector<char> present((size_t)W, 0), is_out((size_t)W, 0);
  for (auto w : N.out_wires) is_out[(size_t)w] = 1;
  for (const auto& ws : N.leaf_wires)
    for (auto a : ws) {
      present[(size_t)a] = 1;
      for (auto b : ws) if (a != b) set_bit(a, b);
    }
  std::vector<int> deg((size_t)W, 0);
  auto degree = [&](int32_t w) {
    int d = 0;
    const uint64_t* r = row(w);
    for (int i = 0; i < NW; i++) d += __builtin_popcountll(r[i]);
    return d;
  };
  for (int w = 0; w < W; w++) deg[(size_t)w] = degree(w);
  std::vector<double> cache((size_t)W, -1.0);
  auto for_each_bit = [&](const uint64_t* r, auto&& f) {
    for (int i = 0; i < NW; i++) {
      uint64_t m = r[i];
      while (m) {
        const int b = __builtin_ctzll(m);
        m &= m - 1;
        f((int32_t)(i * 64 + b));
      }
    }
  };
  auto fill = [&](int32_t w) {
    // pairs of neighbours that are not adjacent = (d * (d - 1) - sum over neighbours a of |adj[a] & nb|) / 2
    const uint64_t* nb = row(w);
    int64_t linked = 0;
    for_each_bit(nb, [&](int32_t a) {
      const uint64_t* ra = row(a);
      for (int i = 0; i < NW; i++) linked += __builtin_popcountll(ra[i] & nb[i]);
    });
    const int64_t d = deg[(size_t)w];
    return (double)((d * (d - 1) - linked) / 2);
  };
  std::vector<int32_t> alive;
  for (int w = 0; w < W; w++) if (present[(size_t)w]) alive.push_back(w);
  std::vector<int32_t> order;
  std::vector<char> dead((size_t)W, 0);
  std::vector<uint64_t> nbcopy((size_t)NW);
  while (true) {
    int32_t best = -1;
    double bk = 0;
    int bd = 0;
    // candidates in increasing wire id: ties resolve like Python's min over (key, degree, wire)
    for (auto w : alive) {
      if (dead[(size_t)w] || is_out[(size_t)w]) continue;
      if (cache[(size_t)w] < 0) cache[(size_t)w] = min_fill ? fill(w) : (double)deg[(size_t)w];
      double k = cache[(size_t)w];
      if (noise > 0.0) k += noise * rng.uniform() * (1.0 + k);
      const int dg = deg[(size_t)w];
      if (best < 0 || k < bk || (k == bk && dg < bd)) { best = w; bk = k; bd = dg; }
    }
    if (best < 0) break;
    std::copy(row(best), row(best) + NW, nbcopy.begin());
    for_each_bit(nbcopy.data(), [&](int32_t a) {
      uint64_t* ra = row(a);
      for (int i = 0; i < NW; i++) ra[i] |= nbcopy[(size_t)i];
      ra[a >> 6] &= ~(uint64_t(1) << (a & 63));
      ra[best >> 6] &= ~(uint64_t(1) << (best & 63));
      deg[(size_t)a] = degree(a);
      cache[(size_t)a] = -1.0;
      for_each_bit(ra, [&](int32_t b) { cache[(size_t)b] = -1.0; });
    });
    dead[(size_t)best] = 1;
    std::fill(row(best), row(best) + NW, 0);
    deg[(size_t)best] = 0;
    order.push_back(best);
    if ((order.size() & 255) == 0) {   // compact the candidate list now and then
      std::vector<int32_t> keep;
      for (auto w : alive) if (!dead[(size_t)w]) keep.push_back(w);
      alive.swap(keep);
    }
  }
  return order;
}

struct TreeBuilder {
  const tsb_network& N;
  std::vector<std::vector<int32_t>> wires_of;  // sorted wire lists per node
  std::vector<int> count;
  std::vector<std::set<int32_t>> holders;
  std::set<int32_t> alive;
  std::vector<int32_t> path;
  int32_t nxt;
  explicit TreeBuilder(const tsb_network& n) : N(n), count((size_t)n.n_wires, 0), holders((size_t)n.n_wires) {
    const int L = (int)N.leaf_wires.size();
    wires_of.resize((size_t)L);
    for (int i = 0; i < L; i++) {
      std::set<int32_t> s(N.leaf_wires[(size_t)i].begin(), N.leaf_wires[(size_t)i].end());
      wires_of[(size_t)i].assign(s.begin(), s.end());
      for (auto w : s) { count[(size_t)w]++; holders[(size_t)w].insert(i); }
      alive.insert(i);
    }
    for (auto w : std::set<int32_t>(N.out_wires.begin(), N.out_wires.end())) count[(size_t)w]++;
    nxt = L;
  }
  static bool has(const std::vector<int32_t>& v, int32_t w) { return std::binary_search(v.begin(), v.end(), w); }
  int out_rank(int a, int b) const {
    const auto &la = wires_of[(size_t)a], &lb = wires_of[(size_t)b];
    int r = 0;
    for (auto w : la) if (count[(size_t)w] - (has(lb, w) ? 1 : 0) >= 2) r++;
    for (auto w : lb) if (!has(la, w) && count[(size_t)w] >= 2) r++;
    return r;
  }
  int merge(int a, int b) {
    const std::vector<int32_t> la = wires_of[(size_t)a], lb = wires_of[(size_t)b];
    std::vector<int32_t> all;
    std::set_union(la.begin(), la.end(), lb.begin(), lb.end(), std::back_inserter(all));
    std::vector<int32_t> o;
    for (auto w : all) {
      const int c = count[(size_t)w] - ((has(la, w) && has(lb, w)) ? 1 : 0);
      holders[(size_t)w].erase(a);
      holders[(size_t)w].erase(b);
      if (c >= 2) {
        o.push_back(w);
        count[(size_t)w] = c;
        holders[(size_t)w].insert(nxt);
      } else {
        count[(size_t)w] = 0;
      }
    }
    wires_of.push_back(o);
    alive.erase(a);
    alive.erase(b);
    alive.insert(nxt);
    path.push_back(a);
    path.push_back(b);
    return nxt++;
  }
};

// Bucket elimination (planner/search.py::_tree_from_order).
std::vector<int32_t> tree_from_order(const tsb_network& N, const std::vector<int32_t>& order) {
  TreeBuilder T(N);
  auto by_size = [&](int a, int b) {
    const size_t sa = T.wires_of[(size_t)a].size(), sb = T.wires_of[(size_t)b].size();
    return sa != sb ? sa < sb : a < b;
  };
  for (auto w : order) {
    std::set<int32_t> bucket(T.holders[(size_t)w].begin(), T.holders[(size_t)w].end());
    while (bucket.size() > 1) {
      std::vector<int32_t> bl(bucket.begin(), bucket.end());
      std::sort(bl.begin(), bl.end(), by_size);
      int br = 0, ba = -1, bb = -1;
      for (size_t i = 0; i < std::min<size_t>(bl.size(), 8); i++)
        for (size_t j = i + 1; j < std::min<size_t>(bl.size(), 9); j++) {
          const int r = T.out_rank(bl[i], bl[j]);
          if (ba < 0 || r < br || (r == br && (bl[i] < ba || (bl[i] == ba && bl[j] < bb)))) { br = r; ba = bl[i]; bb = bl[j]; }
        }
      bucket.erase(ba);
      bucket.erase(bb);
      bucket.insert(T.merge(ba, bb));
    }
  }
  while (T.alive.size() > 1) {
    std::vector<int32_t> rest(T.alive.begin(), T.alive.end());
    std::sort(rest.begin(), rest.end(), by_size);
    T.merge(rest[0], rest[1]);
  }
  return T.path;
}

struct Stats {
  double log2_ops = 0, log2_write = 0, log2_width = 0;
  std::vector<std::vector<int32_t>> step_out;   // output wires of every step (sliced wires removed)
};

// Cost model of planner/tree.py::stats (all wires binary: sizes are wire counts).
Stats tree_stats(const tsb_network& N, const std::vector<int32_t>& ssa, const std::set<int32_t>& sliced) {
  const int L = (int)N.leaf_wires.size();
  std::vector<int> count((size_t)N.n_wires, 0);
  std::vector<std::vector<int32_t>> wires((size_t)(2 * L - 1));
  Stats st;
  for (int i = 0; i < L; i++) {
    std::set<int32_t> s;
    for (auto w : N.leaf_wires[(size_t)i]) if (!sliced.count(w)) s.insert(w);
    wires[(size_t)i].assign(s.begin(), s.end());
    for (auto w : s) count[(size_t)w]++;
    st.log2_width = std::max(st.log2_width, (double)s.size());
  }
  for (auto w : std::set<int32_t>(N.out_wires.begin(), N.out_wires.end())) if (!sliced.count(w)) count[(size_t)w]++;
  double ops = 0, write = 0;
  for (int s = 0; s < L - 1; s++) {
    const auto &la = wires[(size_t)ssa[(size_t)(2 * s)]], &lb = wires[(size_t)ssa[(size_t)(2 * s + 1)]];
    std::vector<int32_t> all, o;
    std::set_union(la.begin(), la.end(), lb.begin(), lb.end(), std::back_inserter(all));
    for (auto w : all) {
      const bool both = std::binary_search(la.begin(), la.end(), w) && std::binary_search(lb.begin(), lb.end(), w);
      const int c = count[(size_t)w] - (both ? 1 : 0);
      if (c >= 2) { o.push_back(w); count[(size_t)w] = c; } else { count[(size_t)w] = 0; }
    }
    ops += std::exp2(1.0 + (double)all.size());
    write += std::exp2((double)o.size());
    st.log2_width = std::max(st.log2_width, (double)o.size());
    wires[(size_t)(L + s)] = o;
    st.step_out.push_back(o);
  }
  const double n_slices = std::exp2((double)sliced.size());
  st.log2_ops = std::log2(std::max(ops, 1.0) * n_slices);
  st.log2_write = std::log2(std::max(write, 1.0) * n_slices);
  return st;
}

double objective(const Stats& st) {  // "combo": ops + 64 * write
  return std::log2(std::exp2(st.log2_ops) + 64.0 * std::exp2(st.log2_write));
}

// planner/tree.py::slice_to_fast: slice the non-output wire that carries the most weight above the target
std::vector<int32_t> slice_to(const tsb_network& N, const std::vector<int32_t>& ssa, double target, int max_sliced) {
  std::vector<int32_t> sliced;
  Stats st = tree_stats(N, ssa, {});
  std::vector<std::set<int32_t>> sets;
  for (auto& o : st.step_out) sets.emplace_back(o.begin(), o.end());
  std::set<int32_t> out(N.out_wires.begin(), N.out_wires.end());
  while (true) {
    std::map<int32_t, double> weight;
    double worst = 0;
    for (auto& ws : sets) {
      const double sz = (double)ws.size();
      worst = std::max(worst, sz);
      if (sz > target + 1e-9) {
        const double bonus = std::exp2(std::min(sz - target, 60.0));
        for (auto w : ws) if (!out.count(w)) weight[w] += bonus;
      }
    }
    if (worst <= target + 1e-9) break;
    TSB_CHECK(!weight.empty() && (int)sliced.size() < max_sliced, TSB_ERR_INVALID,
              "Slicing budget exhausted before reaching the target size.");
    int32_t best = -1;
    double bw = -1;
    for (auto& kv : weight) if (kv.second > bw) { best = kv.first; bw = kv.second; }   // first maximum in wire order
    sliced.push_back(best);
    for (auto& ws : sets) ws.erase(best);
  }
  return sliced;
}

}  // namespace
}  // namespace tsb

using namespace tsb;

extern "C" {

int tsb_network_from_clauses(int32_t n_vars, int32_t n_clauses, const int32_t* clause_sizes, const int32_t* literals,
                             const char* inputs, tsb_network** net) {
  TSB_FE_BEGIN
  TSB_CHECK(net && n_vars >= 0 && n_clauses >= 0 && (n_clauses == 0 || (clause_sizes && literals)), TSB_ERR_INVALID,
            "invalid arguments");
  std::unique_ptr<tsb_network> N(new tsb_network());
  N->n_vars = n_vars;
  const int32_t* p = literals;
  for (int c = 0; c < n_clauses; c++) {
    TSB_CHECK(clause_sizes[c] >= 0, TSB_ERR_INVALID, "negative clause size");
    N->clauses.emplace_back(p, p + clause_sizes[c]);
    for (auto lit : N->clauses.back())
      TSB_CHECK(lit != 0 && std::abs(lit) <= n_vars, TSB_ERR_INVALID, "Clauses contain invalid variables.");
    p += clause_sizes[c];
  }
  if (inputs) {
    check_inputs(*N, inputs);
    N->has_inputs = true;
    N->inputs = inputs;
  }
  build_wiring(*N);
  *net = N.release();
  TSB_FE_END
}

int tsb_network_from_dimacs(const char* text, const char* inputs, tsb_network** net) {
  TSB_FE_BEGIN
  TSB_CHECK(text && net, TSB_ERR_INVALID, "null argument");
  // rows: stripped, non-empty, not starting with 'c' (tensorsat/lib/sat.py:84-87)
  std::vector<std::string> rows;
  {
    std::string s(text);
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    s = b == std::string::npos ? "" : s.substr(b, e - b + 1);
    size_t pos = 0;
    while (pos <= s.size()) {
      size_t nl = s.find('\n', pos);
      if (nl == std::string::npos) nl = s.size();
      std::string ln = s.substr(pos, nl - pos);
      size_t lb = ln.find_first_not_of(" \t\r"), le = ln.find_last_not_of(" \t\r");
      ln = lb == std::string::npos ? "" : ln.substr(lb, le - lb + 1);
      if (!ln.empty() && ln[0] != 'c') rows.push_back(ln);
      pos = nl + 1;
    }
  }
  int nv = 0, nc = 0, used = 0;
  TSB_CHECK(!rows.empty() && std::sscanf(rows[0].c_str(), "p cnf %d %d%n", &nv, &nc, &used) == 2, TSB_ERR_INVALID,
            "DIMACS code must start with 'p cnf <num vars> <num clauses>'.");
  std::vector<int32_t> sizes, lits;
  for (size_t r = 1; r < rows.size(); r++) {
    // literals separated by single spaces; the last token (the terminating 0) is dropped
    std::vector<int32_t> toks;
    size_t pos = 0;
    const std::string& ln = rows[r];
    while (pos <= ln.size()) {
      size_t sp = ln.find(' ', pos);
      if (sp == std::string::npos) sp = ln.size();
      const std::string tok = ln.substr(pos, sp - pos);
      char* endp = nullptr;
      const long v = std::strtol(tok.c_str(), &endp, 10);
      TSB_CHECK(!tok.empty() && endp && *endp == 0, TSB_ERR_INVALID, "invalid literal '%s' in DIMACS clause %zu", tok.c_str(), r);
      toks.push_back((int32_t)v);
      pos = sp + 1;
    }
    toks.pop_back();
    sizes.push_back((int32_t)toks.size());
    lits.insert(lits.end(), toks.begin(), toks.end());
  }
  TSB_CHECK((int)sizes.size() == nc, TSB_ERR_INVALID, "Number of clauses does not match the specified number.");
  return tsb_network_from_clauses(nv, nc, sizes.data(), lits.data(), inputs, net);
  TSB_FE_END
}

int tsb_network_set_inputs(tsb_network* net, const char* inputs) {
  TSB_FE_BEGIN
  TSB_CHECK(net && inputs, TSB_ERR_INVALID, "null argument");
  TSB_CHECK(net->has_inputs, TSB_ERR_INVALID, "this network was built without input boxes");
  check_inputs(*net, inputs);
  net->inputs = inputs;
  for (int v = 0; v < net->n_vars; v++)
    net->leaf_kind[(size_t)v] = inputs[v] == '0' ? TSB_LEAF_BIT_0 : inputs[v] == '1' ? TSB_LEAF_BIT_1 : TSB_LEAF_BIT_UNK;
  TSB_FE_END
}

// Unit propagation of the clauses under a partial assignment: a sound, incomplete filter for the assignment search
// (a partial assignment refuted by propagation has no model; the contraction stays the complete decision).
int tsb_network_propagate(const tsb_network* net, const char* spec, char* out_spec, int32_t* conflict) {
  TSB_FE_BEGIN
  TSB_CHECK(net && spec && conflict, TSB_ERR_INVALID, "null argument");
  check_inputs(*net, spec);
  std::vector<int8_t> val((size_t)net->n_vars);
  for (int v = 0; v < net->n_vars; v++) val[(size_t)v] = spec[v] == '?' ? -1 : (int8_t)(spec[v] - '0');
  *conflict = 0;
  bool changed = true;
  while (changed && !*conflict) {
    changed = false;
    for (const auto& c : net->clauses) {
      int n_free = 0;
      int32_t free_lit = 0;
      bool sat = false;
      for (auto lit : c) {
        const int8_t v = val[(size_t)(std::abs(lit) - 1)];
        if (v < 0) { n_free++; free_lit = lit; }
        else if ((v == 1) == (lit > 0)) { sat = true; break; }
      }
      if (sat) continue;
      if (n_free == 0) { *conflict = 1; break; }
      if (n_free == 1) {
        val[(size_t)(std::abs(free_lit) - 1)] = free_lit > 0 ? 1 : 0;
        changed = true;
      }
    }
  }
  if (out_spec) {
    for (int v = 0; v < net->n_vars; v++) out_spec[v] = val[(size_t)v] < 0 ? '?' : (char)('0' + val[(size_t)v]);
    out_spec[net->n_vars] = 0;
  }
  TSB_FE_END
}

int tsb_network_info(const tsb_network* net, tsb_network_info_t* info) {
  TSB_FE_BEGIN
  TSB_CHECK(net && info, TSB_ERR_INVALID, "null argument");
  info->n_vars = net->n_vars;
  info->n_clauses = (int32_t)net->clauses.size();
  info->n_leaves = (int32_t)net->leaf_wires.size();
  info->n_wires = net->n_wires;
  info->n_out = (int32_t)net->out_wires.size();
  int64_t tot = 0, bytes = 0;
  for (auto& ws : net->leaf_wires) { tot += (int64_t)ws.size(); bytes += int64_t(1) << ws.size(); }
  info->total_leaf_rank = tot;
  info->total_leaf_bytes = bytes;
  TSB_FE_END
}

int tsb_network_arrays(const tsb_network* net, int32_t* leaf_ranks, int32_t* leaf_wires, int32_t* leaf_kinds,
                       int32_t* out_wires) {
  TSB_FE_BEGIN
  TSB_CHECK(net, TSB_ERR_INVALID, "null argument");
  size_t q = 0;
  for (size_t i = 0; i < net->leaf_wires.size(); i++) {
    if (leaf_ranks) leaf_ranks[i] = (int32_t)net->leaf_wires[i].size();
    if (leaf_kinds) leaf_kinds[i] = net->leaf_kind[i];
    if (leaf_wires) for (auto w : net->leaf_wires[i]) leaf_wires[q++] = w;
  }
  if (out_wires) for (size_t i = 0; i < net->out_wires.size(); i++) out_wires[i] = net->out_wires[i];
  TSB_FE_END
}

// the 0/1 tensors of the gates (tensorsat/lib/bincirc.py:42-73), C order, concatenated in slot order
int tsb_network_leaf_tensors(const tsb_network* net, uint8_t* buf, int64_t* offsets) {
  TSB_FE_BEGIN
  TSB_CHECK(net && buf, TSB_ERR_INVALID, "null argument");
  int64_t off = 0;
  for (size_t i = 0; i < net->leaf_kind.size(); i++) {
    if (offsets) offsets[i] = off;
    switch (net->leaf_kind[i]) {
      case TSB_LEAF_BIT_0: buf[off] = 1; buf[off + 1] = 0; off += 2; break;
      case TSB_LEAF_BIT_1: buf[off] = 0; buf[off + 1] = 1; off += 2; break;
      case TSB_LEAF_BIT_UNK: buf[off] = 1; buf[off + 1] = 1; off += 2; break;
      case TSB_LEAF_NOT: {  // [in][out]: out = !in
        const uint8_t t[4] = {0, 1, 1, 0};
        std::memcpy(buf + off, t, 4); off += 4; break;
      }
      default: {            // or_: [a][b][out]: out = a | b
        const uint8_t t[8] = {1, 0, 0, 1, 0, 1, 0, 1};
        std::memcpy(buf + off, t, 8); off += 8; break;
      }
    }
  }
  TSB_FE_END
}

int tsb_network_destroy(tsb_network* net) {
  delete net;
  return TSB_OK;
}

int tsb_tree_search(const tsb_network* net, int32_t trials, uint64_t seed, double target_log2_size, tsb_tree** tree) {
  TSB_FE_BEGIN
  TSB_CHECK(net && tree, TSB_ERR_INVALID, "null argument");
  TSB_CHECK(!net->leaf_wires.empty(), TSB_ERR_INVALID, "Cannot contract an empty network.");
  TSB_CHECK(net->n_wires <= 65536, TSB_ERR_UNSUPPORTED,
            "tree search keeps one adjacency bit row per wire: %d wires is beyond its 65536-wire limit", net->n_wires);
  std::unique_ptr<tsb_tree> T(new tsb_tree());
  T->n_leaves = (int32_t)net->leaf_wires.size();
  if (T->n_leaves > 1) {
    Rng rng(seed);
    const double noises[5] = {0.02, 0.05, 0.1, 0.2, 0.4};
    double best = 0;
    for (int t = 0; t < std::max(1, trials); t++) {
      const bool min_fill = t % 2 == 0;
      const double noise = t < 2 ? 0.0 : noises[rng.next() % 5];
      const std::vector<int32_t> order = elimination_order(*net, min_fill, noise, rng);
      std::vector<int32_t> ssa = tree_from_order(*net, order);
      const Stats st = tree_stats(*net, ssa, {});
      const double obj = objective(st);
      if (t == 0 || obj < best) { best = obj; T->ssa.swap(ssa); }
    }
    T->log2_width_unsliced = tree_stats(*net, T->ssa, {}).log2_width;
    if (target_log2_size > 0 && T->log2_width_unsliced > target_log2_size)
      T->sliced = slice_to(*net, T->ssa, target_log2_size, 4096);
  }
  const Stats st = tree_stats(*net, T->ssa, std::set<int32_t>(T->sliced.begin(), T->sliced.end()));
  T->log2_width = st.log2_width;
  T->log2_ops = st.log2_ops;
  T->log2_write = st.log2_write;
  if (T->n_leaves <= 1) T->log2_width_unsliced = T->log2_width;
  *tree = T.release();
  TSB_FE_END
}

int tsb_tree_info(const tsb_tree* tree, tsb_tree_info_t* info) {
  TSB_FE_BEGIN
  TSB_CHECK(tree && info, TSB_ERR_INVALID, "null argument");
  info->n_leaves = tree->n_leaves;
  info->n_steps = std::max(0, tree->n_leaves - 1);
  info->n_sliced = (int32_t)tree->sliced.size();
  info->log2_width = tree->log2_width;
  info->log2_width_unsliced = tree->log2_width_unsliced;
  info->log2_ops = tree->log2_ops;
  info->log2_write = tree->log2_write;
  TSB_FE_END
}

int tsb_tree_arrays(const tsb_tree* tree, int32_t* ssa_path, int32_t* sliced_wires) {
  TSB_FE_BEGIN
  TSB_CHECK(tree, TSB_ERR_INVALID, "null argument");
  if (ssa_path) std::copy(tree->ssa.begin(), tree->ssa.end(), ssa_path);
  if (sliced_wires) std::copy(tree->sliced.begin(), tree->sliced.end(), sliced_wires);
  TSB_FE_END
}

int tsb_tree_destroy(tsb_tree* tree) {
  delete tree;
  return TSB_OK;
}

int tsb_plan_create_from_tree(tsb_ctx* ctx, const tsb_network* net, const tsb_tree* tree, int32_t mode, int32_t batch_log2,
                              uint32_t flags, tsb_plan** plan) {
  TSB_FE_BEGIN
  TSB_CHECK(ctx && net && tree && plan, TSB_ERR_INVALID, "null argument");
  TSB_CHECK(tree->n_leaves == (int32_t)net->leaf_wires.size(), TSB_ERR_INVALID, "The tree was built for a different network.");
  std::vector<int32_t> ranks, wires, ext((size_t)net->n_wires, 2);
  for (auto& ws : net->leaf_wires) {
    ranks.push_back((int32_t)ws.size());
    wires.insert(wires.end(), ws.begin(), ws.end());
  }
  std::vector<uint8_t> free_((size_t)net->n_wires, 0);   // only the variables multiply a count
  for (int v = 0; v < net->n_vars; v++) free_[(size_t)v] = 1;
  tsb_plan_desc d{};
  d.n_leaves = (int32_t)ranks.size();
  d.leaf_ranks = ranks.data();
  d.leaf_wires = wires.data();
  d.n_out = (int32_t)net->out_wires.size();
  d.out_wires = net->out_wires.data();
  d.n_wires = net->n_wires;
  d.wire_extents = ext.data();
  d.ssa_path = tree->ssa.data();
  d.n_sliced = (int32_t)tree->sliced.size();
  d.sliced_wires = tree->sliced.data();
  d.mode = mode;
  d.wire_is_free = free_.data();
  d.batch_log2 = batch_log2;
  d.flags = flags;
  return tsb_plan_create(ctx, &d, plan);
  TSB_FE_END
}

}  // extern "C"
