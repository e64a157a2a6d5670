// tsb_gemm.cu — batched u8 contraction on the 5th-generation tensor cores (north-star item (c)).
//
//   C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]        A, B: u8, K-major;  accumulate: s32 in TMEM
//   epilogue:  boolean semiring  C = (acc != 0)  -> u8      (FinRel._contract2's np.sign clamp,
//                                                            tensorsat/lang/fin_rel.py:276-279)
//              counting semiring C = acc         -> s32     (#SAT prototype, exact while K*max(A)*max(B) < 2^31)
//
// Pipeline (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d (128B-swizzled boxes of A and B) -> smem ring,
//               completion on mbarrier `full[stage]` (expect_tx)
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::i8 (UMMA 128 x N x 32,
//               operands by shared-memory descriptors, accumulator in TMEM), tcgen05.commit -> `empty[stage]`
//               and, after the last k-block, -> `tmem_full[acc]`
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per load) -> clamp / widen -> 16-byte global stores;
//               two accumulator stages in TMEM (2 x N columns) so that the epilogue of tile i overlaps
//               the MMAs of tile i+1
// Bound: int8 tensor pipe when 2^(m+n+k+1) / (2^(m+k) + 2^(n+k) + e*2^(m+n)) is above the ridge.
#include "lower.h"

#ifndef TSB_HOST_EMU
#include <cuda.h>
#include <cudaTypedefs.h>

namespace tsb {

constexpr int kGemmThreads = 192;
constexpr int kGemmBlockM = 128;
constexpr int kGemmMaxStages = 8;

struct GemmParams {
  int32_t block_n;      // <= 256, power of two >= 32
  int32_t block_k;      // bytes per k-block: 32, 64 or 128 (= swizzle span)
  int32_t num_k_blocks;
  int32_t stages;
  int32_t m_blocks, n_blocks;
  int64_t num_tiles;    // batch * m_blocks * n_blocks
  int64_t n_total;      // N (row length of C)
  int64_t m_total;
  int32_t out_i32;      // 1: exact s32 output, 0: u8
  int32_t clamp;        // u8 output: acc != 0 (else low byte)
  uint32_t idesc;       // tcgen05 instruction descriptor
  uint32_t swizzle_code;  // UMMA layout type (2 = 128B, 4 = 64B, 6 = 32B)
  uint32_t sbo;           // stride byte offset >> 4 (8 rows * block_k bytes)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();  // a broken pipeline must fail loudly, never hang the device
  }
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(sbo & 0x3fff) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell), bits [46,48)
  d |= (uint64_t)(layout_type & 7) << 61;           // swizzle mode, bits [61,64)
  return d;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_i8(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p,
          void* __restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A tile][stages x B tile][barriers][tmem ptr]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = kGemmBlockM * p.block_k;
  const uint32_t b_bytes = p.block_n * p.block_k;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * a_bytes;
  uint64_t* bars = (uint64_t*)(smem_b + (size_t)p.stages * b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kGemmMaxStages;
  uint64_t* tmem_full = bars + 2 * kGemmMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = 2 * p.block_n;  // two accumulator stages; power of two in [64, 512]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  const int64_t tiles_per_batch = (int64_t)p.m_blocks * p.n_blocks;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int batch = (int)(tile / tiles_per_batch);
        const int rem = (int)(tile % tiles_per_batch);
        const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
        for (int kb = 0; kb < p.num_k_blocks; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], a_bytes + b_bytes);
          tma_load_3d(smem_a + (size_t)stage * a_bytes, &map_a, &full[stage], kb * p.block_k, m_blk * kGemmBlockM, batch);
          tma_load_3d(smem_b + (size_t)stage * b_bytes, &map_b, &full[stage], kb * p.block_k, n_blk * p.block_n, batch);
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * p.block_n;
        for (int kb = 0; kb < p.num_k_blocks; kb++) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * a_bytes);
          const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * b_bytes);
          for (int k = 0; k < p.block_k / 32; k++) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, p.sbo, p.swizzle_code);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, p.sbo, p.swizzle_code);
            umma_i8(tmem_d, da, db, p.idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t batch = tile / tiles_per_batch;
      const int rem = (int)(tile % tiles_per_batch);
      const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row = (int64_t)m_blk * kGemmBlockM + quad * 32 + lane;
      const int64_t c_off = (batch * p.m_total + row) * p.n_total + (int64_t)n_blk * p.block_n;
      const uint32_t taddr = tmem_base + acc * p.block_n + ((uint32_t)(quad * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (p.out_i32) {
          uint4* dst = (uint4*)((int32_t*)C + c_off + c0);
#pragma unroll
          for (int q = 0; q < 8; q++) dst[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          uint32_t w[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint32_t e = v[4 * q + j];
              x |= (p.clamp ? (e != 0u ? 1u : 0u) : (e & 0xffu)) << (8 * j);
            }
            w[q] = x;
          }
          uint4* dst = (uint4*)((uint8_t*)C + c_off + c0);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    TSB_CHECK(e == cudaSuccess && p != nullptr, TSB_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

static CUtensorMap make_map(const void* base, uint64_t k, uint64_t rows, uint64_t batch, uint32_t box_k, uint32_t box_rows,
                            CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t dims[3] = {k, rows, batch};
  cuuint64_t strides[2] = {k, k * rows};  // bytes, dims 1 and 2
  cuuint32_t box[3] = {box_k, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSB_CHECK(r == CUDA_SUCCESS, TSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return m;
}

void gemm_u8_impl(tsb_ctx* ctx, const void* A, const void* B, void* C, int b_bits, int m_bits, int n_bits, int k_bits,
                  int out_dtype, int mode, cudaStream_t stream) {
  TSB_CHECK(m_bits >= 7 && n_bits >= 5 && k_bits >= 5, TSB_ERR_UNSUPPORTED,
            "tensor-core contraction needs M >= 128, N >= 32, K >= 32 (got 2^%d, 2^%d, 2^%d)", m_bits, n_bits, k_bits);
  TSB_CHECK(b_bits >= 0 && b_bits + m_bits + k_bits <= 40 && b_bits + n_bits + k_bits <= 40 && k_bits <= 24 && b_bits <= 30,
            TSB_ERR_UNSUPPORTED, "tensor-core contraction: extents out of range");
  TSB_CHECK(out_dtype == TSB_U8 || out_dtype == TSB_I32, TSB_ERR_INVALID, "tensor-core contraction writes u8 or i32");
  TSB_CHECK(((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0), TSB_ERR_INVALID, "operands must be 16-byte aligned");
  GemmParams p{};
  const int64_t M = int64_t(1) << m_bits, N = int64_t(1) << n_bits, K = int64_t(1) << k_bits, Bt = int64_t(1) << b_bits;
  p.block_n = (int)std::min<int64_t>(N, 256);
  p.block_k = (int)std::min<int64_t>(K, 128);
  p.num_k_blocks = (int)(K / p.block_k);
  p.m_blocks = (int)(M / kGemmBlockM);
  p.n_blocks = (int)(N / p.block_n);
  p.num_tiles = Bt * p.m_blocks * p.n_blocks;
  p.n_total = N;
  p.m_total = M;
  p.out_i32 = out_dtype == TSB_I32;
  p.clamp = mode == TSB_SAT;
  const CUtensorMapSwizzle swz = p.block_k == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : p.block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  p.swizzle_code = p.block_k == 128 ? 2u : p.block_k == 64 ? 4u : 6u;
  p.sbo = (uint32_t)(8 * p.block_k) >> 4;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6); a/b format u8 = 0;
  // K-major A and B; N >> 3 at [17,23); M >> 4 at [24,29)
  p.idesc = (2u << 4) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kGemmBlockM >> 4) << 24);
  const size_t stage_bytes = (size_t)kGemmBlockM * p.block_k + (size_t)p.block_n * p.block_k;
  const size_t budget = 200 * 1024;
  p.stages = (int)std::min<size_t>(kGemmMaxStages, std::max<size_t>(2, budget / stage_bytes));
  const size_t smem = 1024 + p.stages * stage_bytes + (2 * kGemmMaxStages + 4) * sizeof(uint64_t) + 16;
  CUtensorMap ma = make_map(A, (uint64_t)K, (uint64_t)M, (uint64_t)Bt, (uint32_t)p.block_k, kGemmBlockM, swz);
  CUtensorMap mb = make_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)Bt, (uint32_t)p.block_k, (uint32_t)p.block_n, swz);
  if (!ctx->gemm_attr_set) {   // per context, hence per device
    TSB_CUDA(cudaFuncSetAttribute(k_gemm_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->gemm_attr_set = true;
  }
  const int grid = (int)std::min<int64_t>(p.num_tiles, ctx->sm_count);
  k_gemm_i8<<<grid, kGemmThreads, smem, stream>>>(ma, mb, p, C);
  ctx->note("k_gemm_i8");
  TSB_CUDA(cudaGetLastError());
}

}  // namespace tsb

extern "C" int tsb_gemm_u8(tsb_ctx* ctx, const void* A, const void* B, void* C, int b_bits, int m_bits, int n_bits,
                           int k_bits, int out_dtype, int mode, tsb_stream stream) {
  try {
    TSB_CHECK(ctx && A && B && C, TSB_ERR_INVALID, "null argument");
    tsb::gemm_u8_impl(ctx, A, B, C, b_bits, m_bits, n_bits, k_bits, out_dtype, mode, (cudaStream_t)stream);
    return TSB_OK;
  } catch (const tsb::Error& e) {
    return e.code;
  }
}

#else  // host emulation build: tcgen05 / TMA cannot be emulated; the entry point reports that

extern "C" int tsb_gemm_u8(tsb_ctx*, const void*, const void*, void*, int, int, int, int, int, int, tsb_stream) {
  tsb::set_error("tsb_gemm_u8 needs an sm_100a device (tcgen05 / TMA are not emulated)");
  return TSB_ERR_UNSUPPORTED;
}

#endif
