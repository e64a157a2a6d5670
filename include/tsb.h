/*
 * tsb.h — C-ABI of libtsb.so, the B200 (sm_100a) contraction executor for TensorSat.
 *
 * Drop-in boundary.  The reference has no native code; each entry point
 * below replaces the numeric body of one reference function (paths relative to the
 * reference repository), and is what a binding for that function would call:
 *
 *   tsb_contract2      FinRel._contract2              tensorsat/lang/fin_rel.py:255-288
 *                      (+ _to_safe_contraction_dtype  tensorsat/lang/fin_rel.py:145-159)
 *   tsb_rewire         FinRel._rewire / rewire_array  tensorsat/lang/fin_rel.py:379-393,
 *                                                     tensorsat/_utils/misc.py:104-123
 *   tsb_permute        np.transpose materialisation   tensorsat/lang/fin_rel.py:381-382
 *   tsb_slice          FinRel._getitem                tensorsat/lang/fin_rel.py:342-343
 *   tsb_union          FinRel._union                  tensorsat/lang/fin_rel.py:345-346
 *   tsb_stack          FinRel._stack                  tensorsat/lang/fin_rel.py:304-308
 *   tsb_spider         FinRel._spider / _scalar       tensorsat/lang/fin_rel.py:290-302
 *   tsb_all_ones       FinRel._as_scalar              tensorsat/lang/fin_rel.py:472-473
 *   tsb_plan_*         ContractionTree.contract as driven by
 *                      CotengraContraction.contract   tensorsat/contractions/cotengra.py:100-128
 *                      and SimpleContraction.contract tensorsat/contractions/simple.py:251-292
 *   tsb_arena_*        per-step numpy allocations     tensorsat/lang/fin_rel.py:310-326
 *
 * Conventions
 *   - Every function returns a tsb_status (0 = ok) and never throws; tsb_last_error()
 *     returns a thread-local message for the last failing call on this thread.
 *   - Tensors are dense, C-ordered (last axis fastest), caller-owned DEVICE buffers
 *     (e.g. torch storage) unless obtained from the arena.  No torch types cross the ABI.
 *   - All work is enqueued on the caller-supplied stream (a cudaStream_t passed as
 *     void*); the only synchronising calls are tsb_all_ones and tsb_scalar_to_host.
 *   - A context is bound to one device.  Its arena is thread-safe and stream-aware; a plan
 *     serialises its executions (an execution on another stream than the previous one waits
 *     for it), so a context may be shared by the streams / threads of one process.
 *   - Wires are small non-negative integers < n_wires; wire_extents[w] is the dimension of
 *     wire w.  Power-of-two extents take the fast kernels, anything else a generic one.
 *   - mode TSB_SAT clamps every result to {0,1} (boolean semiring, count > 0);
 *     TSB_COUNT accumulates exactly (counting semiring).
 */
#ifndef TSB_H_
#define TSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSB_VERSION 100 /* 0.1.0 */
#define TSB_MAX_RANK 62

typedef struct tsb_ctx tsb_ctx;
typedef struct tsb_plan tsb_plan;
typedef void* tsb_stream; /* cudaStream_t */

typedef enum {
  TSB_OK = 0,
  TSB_ERR_INVALID = 1,     /* contract violation (ValueError on the Python side) */
  TSB_ERR_CUDA = 2,        /* CUDA runtime / driver error */
  TSB_ERR_OOM = 3,         /* device allocation failed */
  TSB_ERR_UNSUPPORTED = 4, /* valid but not implemented */
  TSB_ERR_OVERFLOW = 5     /* counting result may not fit in int64 */
} tsb_status;

typedef enum { TSB_U8 = 0, TSB_I32 = 1, TSB_I64 = 2 } tsb_dtype;
typedef enum { TSB_SAT = 0, TSB_COUNT = 1 } tsb_mode;

int tsb_version(void);
const char* tsb_last_error(void);

int tsb_ctx_create(int device, tsb_ctx** ctx);
int tsb_ctx_destroy(tsb_ctx* ctx);
/* Number of kernels launched through this context since creation (bench "gpu_launches"). */
int tsb_ctx_launch_count(tsb_ctx* ctx, int64_t* count);
/* Per-kernel launch counts: tsb_ctx_trace(ctx, 1) clears the counters and starts counting, tsb_ctx_trace_json
 * returns {"kernel name": launches, ...} (names as in DESIGN.md 3.2, e.g. "k_permute<4>", "k_merge_bits_f",
 * "k_gemm_i8").  Tests and smoke() use it to prove which kernels served a call. */
int tsb_ctx_trace(tsb_ctx* ctx, int enabled);
int tsb_ctx_trace_json(tsb_ctx* ctx, char* buf, size_t cap, size_t* needed);

/* ---- caching device allocator for intermediates (north-star item (d)) ------------------- */
typedef struct {
  size_t reserved_bytes;    /* device memory held by the arena */
  size_t in_use_bytes;      /* handed out and not yet released */
  size_t peak_in_use_bytes; /* high-water mark of in_use_bytes */
  size_t device_allocs;     /* number of cudaMalloc calls so far */
  size_t cache_hits;        /* allocations served without touching the driver */
  size_t cross_stream_waits; /* cached blocks handed to another stream than their last user's (ordered by an event) */
} tsb_arena_stats_t;

/* The arena is stream-aware and thread-safe: a block released with tsb_arena_release_async remembers the stream
 * its last user ran on (an event is recorded there), and an allocation for another stream is ordered after it.
 * tsb_arena_alloc / tsb_arena_release are the stream-less forms: the caller guarantees that nothing using a
 * block it releases is still in flight. */
int tsb_arena_reserve(tsb_ctx* ctx, size_t bytes); /* pre-size from a plan's peak estimate */
int tsb_arena_alloc(tsb_ctx* ctx, size_t bytes, void** ptr);
int tsb_arena_release(tsb_ctx* ctx, void* ptr);
int tsb_arena_alloc_async(tsb_ctx* ctx, size_t bytes, tsb_stream stream, void** ptr);
int tsb_arena_release_async(tsb_ctx* ctx, void* ptr, tsb_stream stream);
int tsb_arena_trim(tsb_ctx* ctx); /* return all cached blocks to the driver */
int tsb_arena_stats(tsb_ctx* ctx, tsb_arena_stats_t* stats);

/* ---- single operations (the Box hooks) ---------------------------------------------------- */

/* dst = transpose(src, perm): dst axis i is src axis perm[i] (numpy semantics). */
int tsb_permute(tsb_ctx* ctx, const void* src, void* dst, int rank, const int32_t* extents,
                const int32_t* perm, int elem_size, tsb_stream stream);

/* dst[i_0..i_{out_rank-1}] = sum over dropped src ports of src[...], with dst port j attached
 * to src port out_ports[j]; repeated ports tie dst axes to a diagonal (zero elsewhere). */
int tsb_rewire(tsb_ctx* ctx, const void* src, int src_rank, const int32_t* src_extents,
               int src_dtype, void* dst, int out_rank, const int32_t* out_ports, int dst_dtype,
               int mode, tsb_stream stream);

/* out[out_wires] = sum over wires not in out_wires of lhs[lhs_wires] * rhs[rhs_wires]
 * (einsum with integer labels); out_wires may repeat or omit wires. */
int tsb_contract2(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires,
                  int lhs_dtype, const void* rhs, int rhs_rank, const int32_t* rhs_wires,
                  int rhs_dtype, void* out, int out_rank, const int32_t* out_wires, int out_dtype,
                  int n_wires, const int32_t* wire_extents, int mode, tsb_stream stream);

/* dst = src[fixed] transposed: axis a of src is fixed to fixed_idx[a] (>= 0) or kept (-1);
 * dst axis j is src axis out_axes[j] (the kept axes, in any order). */
int tsb_slice(tsb_ctx* ctx, const void* src, int rank, const int32_t* extents,
              const int32_t* fixed_idx, int out_rank, const int32_t* out_axes, void* dst,
              int elem_size, tsb_stream stream);

/* dst = stack(srcs, axis): n tensors of identical shape, new axis of extent n at `axis`. */
int tsb_stack(tsb_ctx* ctx, const void* const* srcs, int n, int rank, const int32_t* extents,
              int axis, void* dst, int elem_size, tsb_stream stream);

/* out = a | b (TSB_SAT, 0/1 values) or a + b (TSB_COUNT), elementwise over n elements. */
int tsb_union(tsb_ctx* ctx, const void* a, const void* b, void* out, int64_t n, int dtype,
              int mode, tsb_stream stream);

/* delta tensor: num_ports axes of extent dim, ones on the diagonal (num_ports = 0: scalar 1). */
int tsb_spider(tsb_ctx* ctx, void* dst, int dim, int num_ports, int dtype, tsb_stream stream);
int tsb_fill(tsb_ctx* ctx, void* dst, int64_t n, int64_t value, int dtype, tsb_stream stream);
/* dst = (clamp ? src != 0 : src) converted to dst_dtype. */
int tsb_convert(tsb_ctx* ctx, const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
                int clamp, tsb_stream stream);

/* ---- bit-packed boolean relations (1 bit per entry; the reference authors' planned "Binary Packing",
 *      paper/typst/main.typ:336-339) ------------------------------------------------
 * Entry i of a packed tensor is bit (i & 31) of 32-bit word (i >> 5); a packed buffer holds
 * max(16, ceil(n / 8)) bytes and is 16-byte aligned.  Plans use this representation internally for the
 * large steps of a TSB_SAT contraction (tsb_plan_info_t.n_packed_steps); these three entry points expose
 * the same kernels. */
/* dst bit i = (src[i] != 0); src holds n uint8 entries. */
int tsb_pack_bits(tsb_ctx* ctx, const void* src_u8, void* dst_bits, int64_t n, tsb_stream stream);
/* dst[i] = bit i of src (accumulate != 0: dst[i] |= bit i). */
int tsb_unpack_bits(tsb_ctx* ctx, const void* src_bits, void* dst_u8, int64_t n, int accumulate,
                    tsb_stream stream);
/* FinRel._contract2 (tensorsat/lang/fin_rel.py:255-288) with lhs, rhs and out all bit-packed, boolean
 * semiring.  All extents must be 2, the kept wires must appear in the same relative order in lhs, rhs and
 * out with the summed wires leading (the layout plans use), out_rank >= 7 and at most 8 summed wires;
 * anything else returns TSB_ERR_UNSUPPORTED. */
int tsb_contract2_bits(tsb_ctx* ctx, const void* lhs, int lhs_rank, const int32_t* lhs_wires,
                       const void* rhs, int rhs_rank, const int32_t* rhs_wires, void* out, int out_rank,
                       const int32_t* out_wires, int n_wires, const int32_t* wire_extents, tsb_stream stream);

/* result = 1 iff every one of the n elements equals 1.  Synchronises the stream. */
int tsb_all_ones(tsb_ctx* ctx, const void* ptr, int64_t n, int dtype, int* result,
                 tsb_stream stream);
/* result = 1 iff every one of the n elements is 0 or 1 (the value check of FinRel.__new__,
 * tensorsat/lang/fin_rel.py:119-121, for device-resident tensors).  Synchronises the stream. */
int tsb_all_binary(tsb_ctx* ctx, const void* ptr, int64_t n, int dtype, int* result, tsb_stream stream);
/* Reads one element back as int64.  Synchronises the stream. */
int tsb_scalar_to_host(tsb_ctx* ctx, const void* ptr, int dtype, int64_t* out, tsb_stream stream);

/* ---- whole-tree plans --------------------------------------------------------------------- */
typedef struct {
  int32_t n_leaves;
  const int32_t* leaf_ranks;   /* [n_leaves] */
  const int32_t* leaf_wires;   /* concatenated wire ids of every leaf, leaf order */
  int32_t n_out;
  const int32_t* out_wires;    /* wires of the result, in result-axis order (distinct) */
  int32_t n_wires;
  const int32_t* wire_extents; /* [n_wires] */
  const int32_t* ssa_path;     /* [2 * (n_leaves - 1)] pairs; step s creates id n_leaves + s */
  int32_t n_sliced;
  const int32_t* sliced_wires; /* slice id digits, first wire most significant */
  int32_t mode;                /* tsb_mode */
  const uint8_t* wire_is_free; /* optional [n_wires]: 1 if summing the wire can multiply a count
                                  (variables); 0 if it is functionally determined.  NULL = all 1 */
  int32_t batch_log2;          /* tensors whose size incl. sliced wires is <= 2^batch_log2 are
                                  computed once for all slices; < 0 selects the default */
  uint32_t flags;              /* TSB_PLAN_* */
} tsb_plan_desc;

#define TSB_PLAN_NO_GRAPH 1u     /* never replay executions of this plan as a CUDA graph: every call launches its
                                    kernels directly (the default captures the launch sequence on the second
                                    execution with the same slice range / result buffer and replays it afterwards,
                                    DESIGN.md 3.7) */
#define TSB_PLAN_FORCE_GENERIC 2u /* use only the generic kernel (parity reference on device) */
#define TSB_PLAN_NO_GEMM 4u      /* never lower a step to the tcgen05 GEMM */
#define TSB_PLAN_NO_PACK 8u      /* keep every boolean intermediate at one byte per entry (no bit-packing) */
#define TSB_PLAN_NO_FUSE 16u     /* one launch per tree step (no fused n-ary steps) */
/* Environment knobs read by the library (A/B runs and tests; defaults in parentheses):
 *   TSB_PDL (1)              0: plain stream order instead of programmatic dependent launch
 *   TSB_GRAPH (1)            0: no CUDA-graph replay of plan executions (same as TSB_PLAN_NO_GRAPH on every plan)
 *   TSB_GRAPH_STREAMS (4)    side streams used while CAPTURING: the independent launches of a wave of the once-only
 *                            schedule become parallel branches of the graph
 *   TSB_PACK_MIN_LOG2 (20)   steps touching a tensor of at least 2^x entries run bit-packed
 *   TSB_PACK_EXPAND_LOG2 (27) largest broadcast twin of a small operand, log2 entries
 *   TSB_PACK_GATHERED (1)    0: gathered per-slice results are stored as bytes even between packed steps
 *   TSB_FUSE_MIN_LOG2 (18)   consecutive packed steps are fused across intermediates of at least 2^x entries
 *   TSB_FUSE_MAX_K (2)       summed wires per fused launch (-1: never fuse)
 *   TSB_FUSE_PAIR (3)        pair wires of k_merge_bits_f: 0 none, 1 one wire chosen by the fetches that do not depend on
 *                            it, 2 / 3 "private streaming" (up to that many output wires that only ONE operand holds;
 *                            that operand is streamed over their combinations), -1: k_merge_bits_n only
 *   TSB_BITSF_TWO_SIDED (1)  private streaming: private wires taken from a SECOND operand as well (0: none, max 2)
 *   TSB_BITSF_MIN_ITEMS_LOG2 tests only: work items a fused launch keeps when pair wires are taken (default 2^18 for
 *                            private streaming, 2^8 for one pair wire)
 *   TSB_FUSE_LANES (0)       1: k_merge_bits_f picks the wires that index a warp's lanes by a wavefront model
 *                            (measured 2 % slower on the n=50 bench: off)
 *   TSB_CHAIN_MAX_OPS (2)    waves of at most this many small ops join a single-CTA chain launch (0: off)
 *   TSB_WAVE_STREAMS (0)     n > 0: the independent launches of one wave of the once-only schedule are spread over
 *                            the caller's stream and n side streams (fork / join); measured with 4 on the n=50
 *                            bench: device time -1.5 %, end-to-end time +1.6 % (more host work before the first
 *                            kernel of a step): off
 *   TSB_SLICE_LANES (0)      two slice lanes: under graph capture the slices alternate between two parallel branches,
 *                            each with its own copy of the per-slice intermediates (0: for bit-packed plans with at
 *                            least 12 launches per slice, 1: never, 2: wherever the per-slice ops allow it)
 *   TSB_PRE_RESTRICT (1)     0: a rank computes the once-only tensors in full whatever its slice block
 *   TSB_PERMUTE_T4 (1)       0: never use the 4x4 register-transpose permute kernel */

typedef struct {
  int64_t n_slices;
  int64_t peak_bytes;         /* arena bytes needed by one execution */
  int64_t out_numel;
  int32_t out_dtype;          /* u8 for TSB_SAT, i64 for TSB_COUNT */
  int32_t n_steps;
  int32_t n_kernels_shared;   /* kernels run once per execution */
  int32_t n_kernels_per_slice;
  double alg_bytes_shared;    /* algorithmic bytes of the once-only part */
  double alg_bytes_per_slice;
  double alg_ops_shared;
  double alg_ops_per_slice;
  int32_t log2_max_tensor;    /* elements, largest intermediate of one slice */
  int32_t may_overflow;       /* counting bound exceeds int64 */
  double alg_bytes_u8_shared;    /* the same steps at one byte per boolean entry (SURVEY.md §8d's definition); */
  double alg_bytes_u8_per_slice; /* alg_bytes_* count what the plan's launches move (1/8 byte per entry when bit-packed, no intermediates of fused steps) */
  int32_t n_packed_steps;     /* steps executed on bit-packed operands */
  int32_t n_conversions;      /* byte <-> bit conversion launches (per execution of every op once) */
  int32_t n_fused_steps;      /* tree steps folded into the launch of their consumer (their result is never stored) */
  int32_t reserved;
} tsb_plan_info_t;

int tsb_plan_create(tsb_ctx* ctx, const tsb_plan_desc* desc, tsb_plan** plan);
int tsb_plan_info(const tsb_plan* plan, tsb_plan_info_t* info);
/* JSON description of the lowered device program (tests interpret it on the host). */
int tsb_plan_describe(const tsb_plan* plan, char* buf, size_t cap, size_t* needed);
/* Contracts slices [slice_begin, slice_end) and combines them (or / sum) into `out`
 * (out_numel elements of out_dtype, C-order over out_wires).  leaf_ptrs[i] is the u8 0/1
 * device tensor of leaf i in its own axis order. */
int tsb_plan_execute(tsb_plan* plan, const void* const* leaf_ptrs, int64_t slice_begin,
                     int64_t slice_end, void* out, tsb_stream stream);
/* The same contraction on HOST buffers (the call a host-language binding makes per contraction: the reference holds
 * its tensors in host memory, tensorsat/lang/fin_rel.py:328-340).  `host_leaves` holds the u8 0/1 leaves back to back
 * in slot order, each in its own C order (`leaves_bytes` must equal the sum of the leaf sizes of the plan);
 * `host_out` receives out_numel elements of out_dtype (`out_bytes` = out_numel * element size).  One host-to-device
 * copy into a staging buffer the plan owns, the execution, one device-to-host copy, all on `stream`; returns after
 * the result has arrived (the only synchronisation).  Pinned host memory keeps both copies asynchronous. */
int tsb_plan_execute_host(tsb_plan* plan, const void* host_leaves, size_t leaves_bytes, int64_t slice_begin,
                          int64_t slice_end, void* host_out, size_t out_bytes, tsb_stream stream);
/* Timing of executions (device time, CUDA events on the launching stream), reported by tsb_plan_describe:
 *   1: an event pair and a host synchronisation around EVERY launch ("ms" per op; serialises the launches and
 *      removes their overlap: per-op shares for tools, not for totals);
 *   2: four events at the phase boundaries of an execution, no host synchronisation ("phase_ms": once-only
 *      schedule / slice loop / tail + combine, valid once the stream has been synchronised): what bench.py's
 *      roofline uses, the launches inside a phase keep their programmatic overlap;
 *   0: off.  Timed executions launch directly (no CUDA-graph replay). */
int tsb_plan_set_timing(tsb_plan* plan, int enabled);
int tsb_plan_timing_json(tsb_plan* plan, char* buf, size_t cap, size_t* needed);
int tsb_plan_destroy(tsb_plan* plan);

/* ---- front end: DIMACS / clause lists -> flat relation network -> contraction tree -> plan, all on the host and without
 *      Python diagram objects.  Replaces, for CNF inputs, CNFInstance.from_dimacs / inputs / diagram
 *      (tensorsat/lib/sat.py:80-100,165-206,248-262), `>>` + flatten (tensorsat/diagrams.py:387-461,1452-1481) and the
 *      tree search the reference delegates to cotengra (tensorsat/contractions/cotengra.py:78-85).  Wire and slot
 *      numbering are exactly those of `(cnf.inputs(x) >> cnf.diagram()).flatten()`. ------------------------------- */
typedef struct tsb_network tsb_network;
typedef struct tsb_tree tsb_tree;
typedef enum {               /* the gate relations of tensorsat/lib/bincirc.py:42-73 */
  TSB_LEAF_BIT_0 = 0,        /* [1, 0] */
  TSB_LEAF_BIT_1 = 1,        /* [0, 1] */
  TSB_LEAF_BIT_UNK = 2,      /* [1, 1] */
  TSB_LEAF_NOT = 3,          /* [in][out], out = !in */
  TSB_LEAF_OR = 4            /* [a][b][out], out = a | b */
} tsb_leaf_kind;
typedef struct {
  int32_t n_vars, n_clauses, n_leaves, n_wires, n_out;
  int32_t reserved;
  int64_t total_leaf_rank;   /* length of the concatenated leaf wire lists */
  int64_t total_leaf_bytes;  /* bytes of all leaf tensors (one byte per entry) */
} tsb_network_info_t;
typedef struct {
  int32_t n_leaves, n_steps, n_sliced, reserved;
  double log2_width;           /* largest intermediate of one slice, log2 entries */
  double log2_width_unsliced;
  double log2_ops, log2_write; /* whole job (all slices): the cost model cotengra reports */
} tsb_tree_info_t;

/* inputs: NULL = no input boxes (the variables stay open: all-solutions tensor), else one character per variable,
 * '0' / '1' fix it, '?' leaves it unknown (bit_unk box). */
int tsb_network_from_clauses(int32_t n_vars, int32_t n_clauses, const int32_t* clause_sizes, const int32_t* literals,
                             const char* inputs, tsb_network** net);
int tsb_network_from_dimacs(const char* dimacs_text, const char* inputs, tsb_network** net);
/* Swap the input boxes (assignment search: same wiring, same tree, same plan). */
int tsb_network_set_inputs(tsb_network* net, const char* inputs);
/* Unit propagation of the clauses under a partial assignment (spec: '0' / '1' / '?' per variable).  conflict = 1: no
 * model extends it (sound; not complete: conflict = 0 decides nothing).  out_spec (n_vars + 1 chars, optional)
 * receives the assignment with the propagated values filled in.  Host only; used by the assignment search to skip
 * contractions whose answer is implied (notebooks/02-random-ksat.ipynb:707-722 decides every candidate by contraction). */
int tsb_network_propagate(const tsb_network* net, const char* spec, char* out_spec, int32_t* conflict);
int tsb_network_info(const tsb_network* net, tsb_network_info_t* info);
/* Caller-allocated arrays (any may be NULL): leaf_ranks / leaf_kinds [n_leaves], leaf_wires [total_leaf_rank],
 * out_wires [n_out]. */
int tsb_network_arrays(const tsb_network* net, int32_t* leaf_ranks, int32_t* leaf_wires, int32_t* leaf_kinds,
                       int32_t* out_wires);
/* The 0/1 leaf tensors, C order, concatenated in slot order into buf [total_leaf_bytes]; offsets [n_leaves] optional. */
int tsb_network_leaf_tensors(const tsb_network* net, uint8_t* buf, int64_t* offsets);
int tsb_network_destroy(tsb_network* net);
/* Best of `trials` elimination-order trees (min-fill / min-degree, trials >= 2 randomised from `seed`), then sliced
 * until the largest intermediate has at most 2^target_log2_size entries (<= 0: no slicing). */
int tsb_tree_search(const tsb_network* net, int32_t trials, uint64_t seed, double target_log2_size, tsb_tree** tree);
int tsb_tree_info(const tsb_tree* tree, tsb_tree_info_t* info);
int tsb_tree_arrays(const tsb_tree* tree, int32_t* ssa_path /* [2 * n_steps] */, int32_t* sliced_wires /* [n_sliced] */);
int tsb_tree_destroy(tsb_tree* tree);
/* tsb_plan_create on the network and tree above (counting mode: only the variables are free wires). */
int tsb_plan_create_from_tree(tsb_ctx* ctx, const tsb_network* net, const tsb_tree* tree, int32_t mode, int32_t batch_log2,
                              uint32_t flags, tsb_plan** plan);

/* ---- micro-benchmark / test hooks for the individual kernels ------------------------------ */
/* C[b][m][n] = sum_k A[b][m][k] * B[b][n][k], u8 operands, tcgen05 kind::i8 with s32 TMEM
 * accumulators.  out_dtype u8: clamp (acc != 0) when mode is TSB_SAT, else low byte;
 * out_dtype i32: exact.  Extents are given as log2 (bits). */
int tsb_gemm_u8(tsb_ctx* ctx, const void* A, const void* B, void* C, int b_bits, int m_bits,
                int n_bits, int k_bits, int out_dtype, int mode, tsb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TSB_H_ */
