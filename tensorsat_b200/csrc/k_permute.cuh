// k_permute.cuh — bit-permutation transpose for tensors whose extents are all powers of two
// (north-star item (b)).
//
// A rank-r tensor with binary axes is a flat array of 2^r elements and an axis permutation
// is a permutation of the index BITS.  Materialises what the reference leaves as a lazy
// np.transpose view (tensorsat/lang/fin_rel.py:381-382) and numpy later walks with strided
// single-element accesses inside einsum.
//
// One CTA moves one tile: the tile's index bits are the union of the source's low bits and
// the destination's low bits (each >= 128 B contiguous where the tensor allows), so both the
// global reads and the global writes are full-line coalesced 16-byte vectors.  In between the
// tile sits in shared memory in source order behind an XOR bank swizzle; the write phase
// gathers G-byte groups (G = the largest power of two, <= 16, of low index bits the
// permutation leaves in place, including the element size) and packs them into 16-byte
// stores.  Bound: HBM.  Algorithmic bytes: 2 * e * 2^r.
#pragma once
#include "k_merge.cuh"  // BitRuns / deposit

namespace tsb {

constexpr int kPermMaxTileBits = 14;  // 16 KiB tile
constexpr int kPermThreads = 256;
constexpr int kPermMaxVecPerThread = (1 << (kPermMaxTileBits - 4)) / kPermThreads;  // 4

struct PermDesc {
  int32_t g;          // log2 of the gather group in bytes (0..4)
  int32_t tbits;      // tile index bits, in BYTE units (incl. g): tile = 2^tbits bytes
  int32_t a_bits;     // low source bits that are contiguous in the tile (in bytes), >= 4
  int32_t b_bits;     // low destination bits that are contiguous in the tile (in bytes), >= 4
  BitRuns src_hi;     // tile-src index bits [a_bits, tbits) -> source byte offset
  BitRuns dst_hi;     // tile-dst index bits [b_bits, tbits) -> destination byte offset
  BitRuns d2s;        // tile-dst index (bytes) -> tile-src index (bytes), a bit permutation
  BitRuns cta_src;    // CTA index -> source byte offset of the tile
  BitRuns cta_dst;    // CTA index -> destination byte offset of the tile
  int64_t n_tiles;
};

// XOR swizzle of a tile byte address: bank-selecting bits [lo, 7) receive the bits above 7.
TSB_HD uint32_t perm_swz(uint32_t addr, int lo) {
  const uint32_t fold = ((addr >> 7) ^ (addr >> 12)) & 31u;  // 5 bits -> address bits 2..6
  return addr ^ ((fold << 2) & (0x7cu & ~((1u << lo) - 1u)));
}

template <int G>  // gather group bytes
__global__ void __launch_bounds__(kPermThreads)
k_permute(const __grid_constant__ PermDesc d, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  TSB_DYN_SMEM(smem);
  constexpr int g = G == 1 ? 0 : G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : 4;
  constexpr int lo = g < 2 ? 2 : g;  // lowest address bit the swizzle may touch
  const int nvec = 1 << (d.tbits - 4);
  // shared memory: [tile bytes][src_hi table u64][dst_hi table u64][d2s table u16 per vector]
  uint64_t* t_src = (uint64_t*)(smem + (1u << d.tbits));
  uint64_t* t_dst = t_src + (1u << (d.tbits - d.a_bits));
  uint16_t* t_d2s = (uint16_t*)(t_dst + (1u << (d.tbits - d.b_bits)));
  for (int i = threadIdx.x; i < (1 << (d.tbits - d.a_bits)); i += blockDim.x) t_src[i] = deposit(d.src_hi, (uint64_t)i);
  for (int i = threadIdx.x; i < (1 << (d.tbits - d.b_bits)); i += blockDim.x) t_dst[i] = deposit(d.dst_hi, (uint64_t)i);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x)
    t_d2s[i] = (uint16_t)perm_swz((uint32_t)deposit(d.d2s, (uint64_t)i << 4), lo);
  // swizzled tile-src offsets of the G-byte groups inside one 16-byte destination vector
  uint32_t off[16 / G];
#pragma unroll
  for (int e = 0; e < 16 / G; e++) off[e] = perm_swz((uint32_t)deposit(d.d2s, (uint64_t)(e * G)), lo);
  __shared__ RunsSmem runs_s, runs_d;
  runs_to_smem(d.cta_src, runs_s);
  runs_to_smem(d.cta_dst, runs_d);
  __syncthreads();
  const uint32_t amask = (1u << d.a_bits) - 1u, bmask = (1u << d.b_bits) - 1u;
  int64_t tile = blockIdx.x;
  uint64_t off_s = tile < d.n_tiles ? deposit_warp(runs_s, (uint64_t)tile) : 0;
  uint64_t off_d = tile < d.n_tiles ? deposit_warp(runs_d, (uint64_t)tile) : 0;
  // software pipeline: the global loads of tile i+1 are issued (into registers) before the gather / store
  // phase of tile i, so that their latency overlaps with it
  uint4 pre[kPermMaxVecPerThread];
  if (tile < d.n_tiles) {
#pragma unroll
    for (int q = 0; q < kPermMaxVecPerThread; q++) {
      const int v = threadIdx.x + q * kPermThreads;
      if (v < nvec) {
        const uint32_t js = (uint32_t)v << 4;
        pre[q] = *(const uint4*)(src + off_s + ((js & amask) | t_src[js >> d.a_bits]));
      }
    }
  }
  for (; tile < d.n_tiles; tile += gridDim.x) {
    uint8_t* d0 = dst + off_d;
    const bool more = tile + gridDim.x < d.n_tiles;
    if (more) {  // next tile's offsets, one bit run per lane (uniform branch)
      off_s = deposit_warp(runs_s, (uint64_t)(tile + gridDim.x));
      off_d = deposit_warp(runs_d, (uint64_t)(tile + gridDim.x));
    }
    // ---- registers -> shared memory (tile-src order, swizzled)
#pragma unroll
    for (int q = 0; q < kPermMaxVecPerThread; q++) {
      const int v = threadIdx.x + q * kPermThreads;
      if (v < nvec) {
        const uint32_t js = (uint32_t)v << 4;
        uint4 w = pre[q];
        const uint32_t sa = perm_swz(js, lo);
        // the swizzle may flip address bits 2,3 (word within the vector): permute registers instead
        const uint32_t x = ((sa ^ js) >> 2) & 3u;
        if (x & 1u) { uint32_t t = w.x; w.x = w.y; w.y = t; t = w.z; w.z = w.w; w.w = t; }
        if (x & 2u) { uint32_t t = w.x; w.x = w.z; w.z = t; t = w.y; w.y = w.w; w.w = t; }
        *(uint4*)(smem + (sa & ~15u)) = w;
      }
    }
    __syncthreads();
    if (more) {
#pragma unroll
      for (int q = 0; q < kPermMaxVecPerThread; q++) {
        const int v = threadIdx.x + q * kPermThreads;
        if (v < nvec) {
          const uint32_t js = (uint32_t)v << 4;
          pre[q] = *(const uint4*)(src + off_s + ((js & amask) | t_src[js >> d.a_bits]));
        }
      }
    }
    // ---- store phase: tile-dst order, gather groups of G bytes, 16-byte vectors out
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint32_t jd = (uint32_t)v << 4;
      const uint32_t sb = t_d2s[v];
      uint32_t w[4];
      if constexpr (G == 16) {
        uint4 t = *(const uint4*)(smem + sb);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
      } else if constexpr (G == 8) {
        uint2 t0 = *(const uint2*)(smem + sb);
        uint2 t1 = *(const uint2*)(smem + (sb ^ off[1]));
        w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y;
      } else if constexpr (G == 4) {
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = *(const uint32_t*)(smem + (sb ^ off[q]));
      } else if constexpr (G == 2) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t lo16 = *(const uint16_t*)(smem + (sb ^ off[2 * q]));
          const uint32_t hi16 = *(const uint16_t*)(smem + (sb ^ off[2 * q + 1]));
          w[q] = lo16 | (hi16 << 16);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t acc = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) acc |= (uint32_t)smem[sb ^ off[4 * q + j]] << (8 * j);
          w[q] = acc;
        }
      }
      *(uint4*)(d0 + ((jd & bmask) | t_dst[jd >> d.b_bits])) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
  }
}

}  // namespace tsb

namespace tsb {

// ---- byte permutations with no low bit in place: 4x4 byte transposes in registers ---------------
// When the two lowest source bits (a0, a1: the byte inside a source word) are not among the four lowest
// destination bits (c0..c3: the byte inside a destination vector), a thread owns the 64 bytes spanned by
// {a0, a1, c0, c1, c2, c3}: 16 source words (one per c-combination, each holding the 4 a-bytes) are read
// with LDS.32, four 4x4 byte transposes (8 PRMT each) turn them into 16 destination words, and the four
// a-combinations leave as four 16-byte stores.  4 LDS.32 + 8 PRMT per 16 output bytes instead of 16 LDS.U8.
struct PermT4Desc {
  PermDesc p;
  BitRuns u2jd;        // work-item index (tile bits without a0, a1, c0..c3, in destination order) -> tile-dst index
  int32_t a_dst[2];    // destination tile bit of a0, a1
};

TSB_D void transpose4x4_bytes(uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  // in: w_j byte i = M[j][i]; out: w_i byte j = M[j][i]
  const uint32_t t0 = __byte_perm(w0, w1, 0x5140u);  // M00 M10 M01 M11
  const uint32_t t1 = __byte_perm(w2, w3, 0x5140u);  // M20 M30 M21 M31
  const uint32_t t2 = __byte_perm(w0, w1, 0x7362u);  // M02 M12 M03 M13
  const uint32_t t3 = __byte_perm(w2, w3, 0x7362u);  // M22 M32 M23 M33
  w0 = __byte_perm(t0, t1, 0x5410u);                 // M00 M10 M20 M30
  w1 = __byte_perm(t0, t1, 0x7632u);                 // M01 M11 M21 M31
  w2 = __byte_perm(t2, t3, 0x5410u);
  w3 = __byte_perm(t2, t3, 0x7632u);
}

static __global__ void __launch_bounds__(kPermThreads)
k_permute_t4(const __grid_constant__ PermT4Desc dd, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  TSB_DYN_SMEM(smem);
  const PermDesc& d = dd.p;
  constexpr int lo = 2;
  const int nvec = 1 << (d.tbits - 4);
  const int nwork = 1 << (d.tbits - 6);
  uint64_t* t_src = (uint64_t*)(smem + (1u << d.tbits));
  uint64_t* t_dst = t_src + (1u << (d.tbits - d.a_bits));
  uint16_t* t_sb = (uint16_t*)(t_dst + (1u << (d.tbits - d.b_bits)));  // swizzled tile-src address of a work item
  uint16_t* t_jd = t_sb + nwork;                                       // tile-dst index of a work item
  for (int i = threadIdx.x; i < (1 << (d.tbits - d.a_bits)); i += blockDim.x) t_src[i] = deposit(d.src_hi, (uint64_t)i);
  for (int i = threadIdx.x; i < (1 << (d.tbits - d.b_bits)); i += blockDim.x) t_dst[i] = deposit(d.dst_hi, (uint64_t)i);
  for (int i = threadIdx.x; i < nwork; i += blockDim.x) {
    const uint32_t jd = (uint32_t)deposit(dd.u2jd, (uint64_t)i);
    t_jd[i] = (uint16_t)jd;
    t_sb[i] = (uint16_t)perm_swz((uint32_t)deposit(d.d2s, (uint64_t)jd), lo);
  }
  uint32_t off[16];  // swizzled tile-src offsets of the 16 c-combinations
#pragma unroll
  for (int e = 0; e < 16; e++) off[e] = perm_swz((uint32_t)deposit(d.d2s, (uint64_t)e), lo);
  __shared__ RunsSmem runs_s, runs_d;
  runs_to_smem(d.cta_src, runs_s);
  runs_to_smem(d.cta_dst, runs_d);
  __syncthreads();
  const uint32_t amask = (1u << d.a_bits) - 1u, bmask = (1u << d.b_bits) - 1u;
  const uint32_t ja0 = 1u << dd.a_dst[0], ja1 = 1u << dd.a_dst[1];
  int64_t tile = blockIdx.x;
  uint64_t off_s = tile < d.n_tiles ? deposit_warp(runs_s, (uint64_t)tile) : 0;
  uint64_t off_d = tile < d.n_tiles ? deposit_warp(runs_d, (uint64_t)tile) : 0;
  uint4 pre[kPermMaxVecPerThread];
  if (tile < d.n_tiles) {
#pragma unroll
    for (int q = 0; q < kPermMaxVecPerThread; q++) {
      const int v = threadIdx.x + q * kPermThreads;
      if (v < nvec) {
        const uint32_t js = (uint32_t)v << 4;
        pre[q] = *(const uint4*)(src + off_s + ((js & amask) | t_src[js >> d.a_bits]));
      }
    }
  }
  for (; tile < d.n_tiles; tile += gridDim.x) {
    uint8_t* d0 = dst + off_d;
    const bool more = tile + gridDim.x < d.n_tiles;
    if (more) {
      off_s = deposit_warp(runs_s, (uint64_t)(tile + gridDim.x));
      off_d = deposit_warp(runs_d, (uint64_t)(tile + gridDim.x));
    }
#pragma unroll
    for (int q = 0; q < kPermMaxVecPerThread; q++) {
      const int v = threadIdx.x + q * kPermThreads;
      if (v < nvec) {
        const uint32_t js = (uint32_t)v << 4;
        uint4 w = pre[q];
        const uint32_t sa = perm_swz(js, lo);
        const uint32_t x = ((sa ^ js) >> 2) & 3u;
        if (x & 1u) { uint32_t t = w.x; w.x = w.y; w.y = t; t = w.z; w.z = w.w; w.w = t; }
        if (x & 2u) { uint32_t t = w.x; w.x = w.z; w.z = t; t = w.y; w.y = w.w; w.w = t; }
        *(uint4*)(smem + (sa & ~15u)) = w;
      }
    }
    __syncthreads();
    if (more) {
#pragma unroll
      for (int q = 0; q < kPermMaxVecPerThread; q++) {
        const int v = threadIdx.x + q * kPermThreads;
        if (v < nvec) {
          const uint32_t js = (uint32_t)v << 4;
          pre[q] = *(const uint4*)(src + off_s + ((js & amask) | t_src[js >> d.a_bits]));
        }
      }
    }
    for (int u = threadIdx.x; u < nwork; u += blockDim.x) {
      const uint32_t sb = t_sb[u] & ~3u;  // word address (the a-bits are the byte lanes)
      uint32_t W[16];
#pragma unroll
      for (int e = 0; e < 16; e++) W[e] = *(const uint32_t*)(smem + (sb ^ off[e]));
      // W[c3c2c1c0] byte a  ->  per (c3,c2): transpose over (c1,c0) x a
#pragma unroll
      for (int q = 0; q < 4; q++) transpose4x4_bytes(W[4 * q], W[4 * q + 1], W[4 * q + 2], W[4 * q + 3]);
      // now W[4*q + a] holds bytes c1c0 = 0..3 for (c3,c2) = q and a-combination a
      const uint32_t jd = t_jd[u];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const uint32_t j = jd | ((a & 1) ? ja0 : 0u) | ((a & 2) ? ja1 : 0u);
        *(uint4*)(d0 + ((j & bmask) | t_dst[j >> d.b_bits])) = make_uint4(W[a], W[4 + a], W[8 + a], W[12 + a]);
      }
    }
    __syncthreads();
  }
}

}  // namespace tsb
