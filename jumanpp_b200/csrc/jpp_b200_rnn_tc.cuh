// jpp_b200_rnn_tc.cuh -- the RNN level step on the 5th-generation tensor cores.
//
// One level of the prefix tree is a dense contraction: ctx'[n][i] = sigmoid(sum_k W[i][k] ctx[prev(n)][k] +
// emb[id_n][i]) for the ~50 000 tree nodes n of the level (reference mikolov_rnn_impl.h:202-214, batched
// over sentences here) -- D[128 x N] = W[128 x 128] . C^T[128 x N].  k_rnn_level_x runs it on the FP32
// pipes in the float order of oracle/eigen_shim (32 residue partials, bit-pinned to the oracle);
// this kernel runs it as tcgen05.mma (kind::tf32, FP32 accumulator in TMEM) with the 3xTF32 split
//     W = Wh + Wl,  c = ch + cl  (h = top 19 bits, l = the exact remainder)
//     W.c ~= Wh.ch + Wl.ch + Wh.cl            (the dropped Wl.cl is 2^-22 of the product)
// so the result carries about as many good bits as the FP32 path, in a different association: it is
// held to the contract of BASELINE.json (RNN scores within 1e-4 relative, 1-best identical), not to
// bit equality, and can be switched off (JPP_B200_RNN_TC=0).  Layer size 128 only (the bench model);
// other sizes take k_rnn_level_x.
//
// Block = kTcGroups groups of 5 warps, each group a pipeline of its own over alternate tiles (own B planes,
// own mbarrier, own TMEM columns, named barriers instead of __syncthreads): while one group waits for its
// MMAs another stages or reads back.  Inside a group, warps 0-3: stage the previous contexts of a tile of
// kTcTile tree nodes into shared memory (split into the two TF32 planes, K-major, 128-byte swizzle: the
// canonical UMMA operand layout), and, after the MMAs, read the accumulator back (tcgen05.ld, thread =
// output i = TMEM lane) for the epilogue: + embedding, sigmoid, coalesced store of the new context rows.
// Warp 4, one elected lane: issues the 48 tcgen05.mma of a tile and commits them to an mbarrier.  The
// group is software-pipelined: the row records and the context rows of tile t+1 are requested (into
// registers) before the group waits for the MMAs of tile t, so the global-memory latencies of a tile hide
// behind the tensor-core work and the epilogue of the one before.  Which rows a tile reads and writes comes
// from one 16-byte record per tree node that k_rnn_level_fill leaves in level order (TcRec) -- not from
// the four dependent index loads the FFMA kernel does.  W (both planes, 128 KB, prepared in the
// shared-memory image once per model) stays in shared memory for the launch.
#pragma once

namespace {

using namespace jppb200;

constexpr u32 kTcH = 128;          // layer size this kernel is built for
#ifndef JPP_TC_TILE
#define JPP_TC_TILE 16
#endif
#ifndef JPP_TC_GROUPS
#define JPP_TC_GROUPS 4
#endif
constexpr u32 kTcTile = JPP_TC_TILE;      // tree nodes per tile (UMMA N): 16 or 32
constexpr u32 kTcGroups = JPP_TC_GROUPS;  // independent tile pipelines per block (each: 4 producer / epilogue warps + 1 MMA warp).
// Four, not six (contexts 6.7 against 6.0 ms alone): 193 KB of shared memory instead of 226 let the block share an SM with
// the walk blocks of another window, which is worth more end to end (jpp_b200.cu, JPP_CUTOFF_SM_R)
constexpr u32 kTcGroupThreads = 160;
constexpr u32 kTcThreads = kTcGroups * kTcGroupThreads;
constexpr u32 kTcChunkK = 32;      // tf32 elements per 128-byte swizzle row
constexpr u32 kTcABytes = kTcH * kTcH * 4;        // one plane of W
constexpr u32 kTcBBytes = kTcTile * kTcH * 4;     // one plane of a context tile
constexpr u32 kTcTmemCols = kTcGroups * kTcTile <= 32 ? 32 : (kTcGroups * kTcTile <= 64 ? 64 : 128);  // power of two >= 32
constexpr u32 kTcSmem = 2 * kTcABytes + kTcGroups * 2 * kTcBBytes + 1024 /* alignment */ + 64;

// byte offset of element (row r, k) of a K-major operand with `rows` rows in the UMMA SWIZZLE_128B layout:
// per 32-element K chunk a region of rows x 128 B; inside, 8-row groups of 1024 B, the 16-byte column of a
// row XORed with (row % 8)
__host__ __device__ inline u32 tc_operand_offset(u32 rows, u32 r, u32 k) {
  const u32 kc = k / kTcChunkK, kk = k % kTcChunkK;
  const u32 c16 = (kk * 4) / 16, within = (kk * 4) % 16;
  return kc * rows * 128 + (r / 8) * 1024 + (r % 8) * 128 + ((c16 ^ (r % 8)) * 16) + within;
}

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61.  K-major + swizzle: LBO is not used (1), SBO = 1024 B
// between 8-row groups.
__device__ __forceinline__ u64 tc_smem_desc(u32 saddr) {
  return (u64)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
constexpr u32 kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTcTile >> 3) << 17) | ((kTcH >> 4) << 24);

__device__ __forceinline__ void tc_mma(u32 tmemD, u64 descA, u64 descB, u32 accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(kTcIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// one tree node of a level: context rows read / written and the embedding row (k_rnn_level_fill)
struct __align__(16) TcRec {
  u32 inRow, outRow, embRow, rnode;
};

__global__ void __launch_bounds__(kTcThreads, 1)
    k_rnn_level_tc(RnnWork rw, RnnTables rt, const float* __restrict__ wImage /* [2][H*H] planes in the smem image */,
                   const TcRec* __restrict__ levelRec, u32 level) {
  extern __shared__ unsigned char tcSmemRaw[];
  __shared__ __align__(8) unsigned long long sBar[kTcGroups];
  __shared__ u32 sTmem;
  __shared__ u32 sInRow[kTcGroups][2][kTcTile];  // rows a tile reads (this tile / the group's next one)
  __shared__ u32 sOutRow[kTcGroups][kTcTile], sEmbRow[kTcGroups][kTcTile];
  const u32 cnt = rw.levelCount[level];
  const u32 tiles = (cnt + kTcTile - 1) / kTcTile;
  if (blockIdx.x * kTcGroups >= tiles) return;
  // operands need 1024-byte alignment (the swizzle is a function of the address bits)
  unsigned char* base = tcSmemRaw + ((1024 - (smem_u32(tcSmemRaw) & 1023)) & 1023);
  unsigned char* sA = base;                    // [2 planes][H x H]
  const u32 group = threadIdx.x / kTcGroupThreads, gtid = threadIdx.x % kTcGroupThreads;
  const u32 warp = gtid >> 5, lane = threadIdx.x & 31u;       // warp inside the group: 0..3 stage / read back, 4 issues
  unsigned char* sB = base + 2 * kTcABytes + group * 2 * kTcBBytes;    // this group's [2 planes][tile x H]
  const TcRec* recs = levelRec + rw.levelOff[level];
  const u32 barId = 1 + group;                 // named barrier of the group (0 is __syncthreads)
  const u32 prodBar = 1 + kTcGroups + group;   // ... of its four staging warps

  // one-time: W planes into shared memory, the mbarriers, the TMEM accumulators (kTcTile columns per group)
  {
    const uint4* src = reinterpret_cast<const uint4*>(wImage);
    uint4* dst = reinterpret_cast<uint4*>(sA);
    for (u32 i = threadIdx.x; i < 2 * kTcABytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    for (u32 g = 0; g < kTcGroups; ++g) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[g])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x / 32 == 4) {  // (the MMA warp of group 0)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTmem)), "r"(kTcTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy writes of W become visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmemBase = sTmem;
  const u32 tmemD = tmemBase + group * kTcTile;
  u32 phase = 0, buf = 0;
  // a warp reads the TMEM lanes of its quarter: (warp id in the CTA) % 4
  const u32 quarter = (threadIdx.x / 32) & 3u;
  const u32 stride = gridDim.x * kTcGroups;
  u32 tile = blockIdx.x * kTcGroups + group;

  // the record of tree node j of a tile is held by thread j of the group: the row it reads goes to the row
  // list at once (the rows are requested a tile ahead), the rows of the epilogue when that tile is current
  auto fetch_record = [&](u32 t) -> uint4 {
    const u32 first = t * kTcTile;
    const u32 j = first + gtid < cnt ? gtid : 0u;   // the tail of the last tile repeats its first node (never stored)
    return __ldg(reinterpret_cast<const uint4*>(recs + first + j));
  };
  uint4 rec = make_uint4(0u, 0u, 0u, 0u);
  float4 v[kTcTile / 4];  // the context rows of the tile about to be staged: row warp + 4 q, columns 4 lane .. 4 lane + 3
  if (tile < tiles && warp < 4) {
    if (gtid < kTcTile) {
      rec = fetch_record(tile);
      sInRow[group][0][gtid] = rec.x;
    }
    asm volatile("bar.sync %0, 128;" ::"r"(prodBar) : "memory");
#pragma unroll
    for (u32 q = 0; q < kTcTile / 4; ++q)
      v[q] = __ldg(reinterpret_cast<const float4*>(rw.ctx + (size_t)sInRow[group][0][warp + 4 * q] * kTcH) + lane);
  }

  for (; tile < tiles; tile += stride) {
    const u32 first = tile * kTcTile;
    const u32 nvalid = (cnt - first) < kTcTile ? (cnt - first) : kTcTile;
    if (warp < 4) {
      // previous contexts -> the two TF32 planes of the B operand (row = tree node, K-major)
#pragma unroll
      for (u32 q = 0; q < kTcTile / 4; ++q) {
        const u32 r = warp + 4 * q;
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(v[q].x) & 0xFFFFE000u); lo.x = __fsub_rn(v[q].x, hi.x);
        hi.y = __uint_as_float(__float_as_uint(v[q].y) & 0xFFFFE000u); lo.y = __fsub_rn(v[q].y, hi.y);
        hi.z = __uint_as_float(__float_as_uint(v[q].z) & 0xFFFFE000u); lo.z = __fsub_rn(v[q].z, hi.z);
        hi.w = __uint_as_float(__float_as_uint(v[q].w) & 0xFFFFE000u); lo.w = __fsub_rn(v[q].w, hi.w);
        const u32 o = tc_operand_offset(kTcTile, r, 4 * lane);
        *reinterpret_cast<float4*>(sB + o) = hi;
        *reinterpret_cast<float4*>(sB + kTcBBytes + o) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // B is staged, and the accumulator of the previous tile has been read back by all four warps
    asm volatile("bar.sync %0, %1;" ::"r"(barId), "r"(kTcGroupThreads) : "memory");
    if (warp == 4) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const u32 aBase = smem_u32(sA), bBase = smem_u32(sB);
        u32 acc = 0;
        // small terms first: Wl.ch, Wh.cl, then Wh.ch
#pragma unroll 1
        for (u32 term = 0; term < 3; ++term) {
          const u32 aPlane = term == 0 ? 1u : 0u, bPlane = term == 1 ? 1u : 0u;
          for (u32 kc = 0; kc < kTcH / kTcChunkK; ++kc) {
#pragma unroll
            for (u32 ks = 0; ks < kTcChunkK / 8; ++ks) {  // UMMA K = 8 for tf32: 32 bytes along the swizzled row
              const u64 da = tc_smem_desc(aBase + aPlane * kTcABytes + kc * kTcH * 128 + ks * 32);
              const u64 db = tc_smem_desc(bBase + bPlane * kTcBBytes + kc * kTcTile * 128 + ks * 32);
              tc_mma(tmemD, da, db, acc);
              acc = 1;
            }
          }
        }
        // arrives on the barrier when every MMA above has completed (implies fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sBar[group])) : "memory");
      }
      __syncwarp();
    } else {
      // while the tensor core works: the rows of this tile's epilogue, the record and the context rows of the
      // group's next tile
      const u32 next = tile + stride;
      if (gtid < kTcTile) {
        sOutRow[group][gtid] = rec.y;
        sEmbRow[group][gtid] = rec.z;
        if (next < tiles) {
          rec = fetch_record(next);
          sInRow[group][buf ^ 1u][gtid] = rec.x;
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(prodBar) : "memory");
      float4 vn[kTcTile / 4] = {};
      if (next < tiles) {
#pragma unroll
        for (u32 q = 0; q < kTcTile / 4; ++q)
          vn[q] = __ldg(reinterpret_cast<const float4*>(rw.ctx + (size_t)sInRow[group][buf ^ 1u][warp + 4 * q] * kTcH) + lane);
      }
      // thread = output i = TMEM lane; the group's columns = the tile's tree nodes
      const u32 i = quarter * 32u + lane;
      const u32 taddr = tmemD + ((quarter * 32u) << 16);
      bool waited = false;
#pragma unroll
      for (u32 c0 = 0; c0 < kTcTile; c0 += 8) {
        // the embedding values of this chunk are requested first (they do not depend on the accumulator)
        float e[8];
#pragma unroll
        for (u32 j = 0; j < 8; ++j) e[j] = __ldg(rt.emb + (size_t)sEmbRow[group][c0 + j] * kTcH + i);
        if (!waited) {
          mbar_wait(smem_u32(&sBar[group]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          waited = true;
        }
        u32 a[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                     : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (u32 j = 0; j < 8; ++j) {
          const float x = __fadd_rn(__uint_as_float(a[j]), e[j]);
          const float r = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
          if (c0 + j < nvalid) rw.ctx[(size_t)sOutRow[group][c0 + j] * kTcH + i] = r;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
      for (u32 q = 0; q < kTcTile / 4; ++q) v[q] = vn[q];
    }
    phase ^= 1u;
    buf ^= 1u;
  }
  __syncthreads();
  if (threadIdx.x / 32 == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(kTcTmemCols));
  }
}

}  // namespace
