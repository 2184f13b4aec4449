// tcgen05 3xTF32 GEMM engine for the dense Kronecker factors (sm_100a).
//
//   C = epilogue( op(A) op(B)  -  op(A2) op(B2) )          float32 in, float32-faithful out
//
// Replaces the large tf.matmul call sites of psgd.py (:173, :175-176, :179, :190, :192, :220, :243, :246, :261,
// :263, :295, :301, :307, :319, :321) for layers big enough to fill 128-wide tensor-core tiles.
//
// Kernel anatomy (one persistent CTA per SM, 10 warps, warp-specialised):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D (SWIZZLE_128B) of raw fp32 operand tiles -> shared memory
//   warps 6-9   splitter       hi = rna_tf32(x), lo = rna_tf32(x - hi): the 3xTF32 split is done on chip, so operands
//                              are read from HBM/L2 exactly once, as plain fp32.  B is split in shared memory (hi in
//                              place, lo in a second tile).  A is split in REGISTERS and stored to TENSOR MEMORY
//                              (tcgen05.st; one thread per A row = TMEM lane), so the tensor core reads A from TMEM
//                              ("TS" mode) and A never costs shared-memory bandwidth again -- the kernel is bound by
//                              shared-memory bandwidth (ncu: profiles/), and this takes the per-stage traffic from
//                              224 KB to 144 KB.  An MN-major A (transpose_a) is transposed by the same register pass.
//   warp 1      MMA issuer     one lane issues tcgen05.mma.kind::tf32: lo*hi + hi*lo + hi*hi per K=8 atom, fp32
//                              accumulators in TMEM (double buffered: 2 x BN columns)
//   warps 2-5   epilogue       tcgen05.ld -> registers -> triu mask / max|.| / (D - mu*acc) / column scale -> global
// Pipelines: full[s] (TMA -> splitter), conv[s] (splitter -> MMA), empty[s] (tcgen05.commit -> TMA),
//            tmem_full[a] / tmem_empty[a] (MMA <-> epilogue).
// Both operands may be K-major or MN-major (tf32 supports both), so every transposition the reference asks for
// (transpose_a / transpose_b) is a descriptor flag, never a copy.  Triangular operands restrict each tile's K range
// and triu outputs skip tiles below the diagonal.
#include <cuda.h>
#include <stdlib.h>

#include <vector>

#include "gemm_tc.cuh"

namespace psgd {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per stage along K = one 128-byte swizzle span
constexpr int UMMA_K = 8;              // tf32
constexpr int kThreads = 320;          // 10 warps
constexpr int kSplitWarp0 = 6;         // warps 2-5: epilogue, 6-9: splitter
constexpr int kChunkKB = 4;            // K-blocks (of 32) per tensor-core accumulation chain
constexpr int kEpiLd = 36;             // row pitch (floats) of an epilogue warp's 32 x 32 shared-memory stage: 16-byte
                                       // accesses by row (lane = row) and by row segment (8 lanes = one row) are both conflict-free

// TS = true : A operand through tensor memory.  Stage = [A raw | B hi | B lo]; TMEM = 2 accumulators (2 x BN columns)
//             + per stage 32 columns of A hi and 32 of A lo (BK = 32 tf32 values per row).
// TS = false: both operands from shared memory ("SS").  Stage = [A hi | B hi | A lo | B lo].
template <int BN, bool TS>
struct Cfg {
  static constexpr int kStages = TS ? 4 : ((BN == 256) ? 2 : 3);
  static constexpr int kTileABytes = BM * BK * 4;             // 16 KB
  static constexpr int kTileBBytes = BN * BK * 4;
  static constexpr int kStageBytes = TS ? (kTileABytes + 2 * kTileBBytes) : 2 * (kTileABytes + kTileBBytes);
  static constexpr int kTxBytes = kTileABytes + kTileBBytes;
  static constexpr int kBHiOff = kTileABytes;
  static constexpr int kBLoOff = TS ? (kTileABytes + kTileBBytes) : (2 * kTileABytes + kTileBBytes);
  static constexpr int kALoOff = kTileABytes + kTileBBytes;   // SS only
  static constexpr int kAccCols = 2 * BN;
  static constexpr int kATmemCol0 = kAccCols;                 // TS only: first column of the A stages
  static constexpr int kTmemCols = TS ? 512 : 2 * BN;
  static_assert(!TS || kAccCols + kStages * 2 * BK <= 512, "TMEM budget");
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                       4 * 32 * kEpiLd * 4 /*epilogue transposition stages*/;
};

// One launch covers up to kMaxGroup problems of identical shape and flags ("batched launch over all layers"): a tile
// index decodes to (problem, tile); only pointers and tensor maps differ per problem.
constexpr int kMaxGroup = 24;

struct GroupMaps {
  CUtensorMap a0[kMaxGroup], b0[kMaxGroup], a1[kMaxGroup], b1[kMaxGroup];
};

struct Params {
  int M, N;
  int K[2];                 // K of product 0 and of the (subtracted) product 1; K[1] == 0 when absent
  int a_mn[2], b_mn[2];     // 1: operand is MN-major in memory (op = transpose of the row-major array)
  int a_tri, b_tri;         // K-range hints for product 0: 0 none, 1 op(A)/op(B) upper triangular, 2 lower
  int ldc, ldd;
  int triu;
  float step, tiny;
  int colscale_recip, colscale_sq;
  int tiles_m, tiles_n, count;
  int pair_b, pair_kind;    // block-pair mode (triangular block-inverse doubling): only tiles with (m0/b even, n0/b == m0/b + 1)
                            // exist; K range = the n-block (kind 1) or the m-block (kind 2)
  int negate;               // C = -(acc)
  int d_tri;                // D is an upper-triangular factor that the b_full flag of the problem vouches for
  int epi;                  // epilogue stores: 2 staged through shared memory (full lines), 1 256-bit stores, 0 row-wise 128-bit
  int debug;                // timing ablations only (wrong results): 1 skip the B split, 2 skip the A split, 4 skip the MMAs,
                            // 8 skip the epilogue's global stores
  float* C[kMaxGroup];
  float* maxabs[kMaxGroup];
  const float* D[kMaxGroup];
  const float* mu_max[kMaxGroup];
  const float* colscale[kMaxGroup];
  const int* a_full[kMaxGroup];   // run-time "operand is not triangular after all" flags (nullptr: hint valid)
  const int* b_full[kMaxGroup];
  const float* rho[kMaxGroup];    // rho_mode 1: C / *rho, 2: C * *rho
  int rho_mode;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// One lane of a fully converged warp.  Guarding the single-thread tcgen05 / TMA instructions with elect.sync (rather than
// `lane == 0`) lets ptxas know exactly one lane is active, so their uniform-register operands need no per-lane
// serialisation loop (with `lane == 0` every UTCHMMA was wrapped in an ELECT/BRA.U.ANY loop, ~20 extra instructions).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is [128 lanes = rows] x [8 columns = K values] of 32-bit tf32, K-major only
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive columns (thread i of the warp owns lane base+i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout, version 1).
//   K-major : rows of 128 B, 8-row groups 1024 B apart            -> LBO = 1 (unused), SBO = 1024 B
//   MN-major: 32-float (128 B) runs along M/N, 8 K-rows per 1024 B atom, next 32-wide run `lbo_bytes` away
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major operand tile: TMA SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart.
// MN-major operand tile (32-bit elements): the only legal layout is SWIZZLE_128B_BASE32B (TMA
// SWIZZLE_128B_ATOM_32B): 128-byte runs along M/N, 4 K-rows per 512-byte swizzle atom; our tile stores each
// 32-wide M/N run as a contiguous [32 k x 128 B] block of 4096 B.
__device__ __forceinline__ uint64_t operand_desc(uint32_t tile_addr, int mn_major, int kk) {
  return mn_major ? make_smem_desc(tile_addr + kk * 1024, /*LBO*/ 4096, /*SBO*/ 512, /*SW128_BASE32B*/ 1)
                  : make_smem_desc(tile_addr + kk * 32, /*LBO*/ 16, /*SBO*/ 1024, /*SW128*/ 2);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32.
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn, int negate_a) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format  = F32
  d |= 2u << 7;                        // a_format  = TF32
  d |= 2u << 10;                       // b_format  = TF32
  d |= (uint32_t)(negate_a & 1) << 13; // a_negate
  d |= (uint32_t)(a_mn & 1) << 15;     // a_major   (1 = MN-major)
  d |= (uint32_t)(b_mn & 1) << 16;     // b_major
  d |= (uint32_t)(n >> 3) << 17;       // n_dim
  d |= (uint32_t)(BM >> 4) << 24;      // m_dim
  return d;
}

// 3xTF32 split of one fp32 value: hi = x rounded to tf32 (nearest, ties away from zero -- exactly cvt.rna.tf32.f32 for
// finite values: add half an ulp of the 10-bit mantissa to the sign-magnitude bit pattern and drop the low 13 bits;
// two integer ops where ptxas expands the cvt to four with its Inf/NaN select), lo = x - hi (exact in fp32).  lo is
// handed to the tensor core unrounded: kind::tf32 ignores the low 13 mantissa bits of its operands, which truncates
// lo at 2^-21 |x| -- below the dropped lo*lo term.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = (x + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(__uint_as_float(x) - __uint_as_float(hi));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// in-place hi + second-tile lo for `bytes` of raw fp32 in shared memory, 128 threads, conflict-free 16-byte accesses
__device__ __forceinline__ void split_tile_smem(uint32_t hi_addr, uint32_t lo_addr, int bytes, int st) {
#pragma unroll 4
  for (int off = st * 16; off < bytes; off += 128 * 16) {
    const uint4 x = lds128(hi_addr + off);
    uint4 h, l;
    split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
    sts128(hi_addr + off, h);
    sts128(lo_addr + off, l);
  }
}

// Does output tile (m0, n0) take part in the product at all?  (tiles skipped by triu are still WRITTEN, as zeros, by the
// epilogue; tiles outside the block-pair pattern are not touched.)
template <int BN>
__host__ __device__ __forceinline__ bool tile_in_pattern(const Params& p, int m0, int n0) {
  if (p.pair_b) {
    const int bm = m0 / p.pair_b;
    if ((bm & 1) || n0 / p.pair_b != bm + 1) return false;
  }
  return true;
}
template <int BN>
__host__ __device__ __forceinline__ bool tile_computed(const Params& p, int m0, int n0) {
  return tile_in_pattern<BN>(p, m0, n0) && !(p.triu && m0 >= n0 + BN);
}

// K-block range [kb0, kb1) of output tile (m0, n0) (BMt rows x BN columns) for product `p`
template <int BN, int BMt = BM>
__host__ __device__ __forceinline__ void k_range(const Params& p, int prod, int m0, int n0, int& kb0, int& kb1,
                                                 int full = 0) {
  const int K = p.K[prod];
  int lo = 0, hi = K;
  auto up = [](int& v, int x) { if (x > v) v = x; };
  auto down = [](int& v, int x) { if (x < v) v = x; };
  if (prod == 0) {
    // full: bit 0 = drop the hint on A, bit 1 = drop the hint on B (the factor turned out not to be triangular)
    const int a_tri = (full & 1) ? 0 : p.a_tri, b_tri = (full & 2) ? 0 : p.b_tri;
    if (a_tri == 1) up(lo, m0);                         // op(A)[m,k] = 0 for k < m
    if (a_tri == 2) down(hi, m0 + BMt);                 // op(A)[m,k] = 0 for k > m
    if (b_tri == 1) down(hi, n0 + BN);                  // op(B)[k,n] = 0 for k > n
    if (b_tri == 2) up(lo, n0);                         // op(B)[k,n] = 0 for k < n
    if (p.pair_b) {
      const int base = (p.pair_kind == 1 ? n0 : m0) / p.pair_b * p.pair_b;
      up(lo, base);
      down(hi, base + p.pair_b);
    }
  }
  kb0 = lo / BK;
  kb1 = (hi + BK - 1) / BK;
  if (kb1 < kb0) kb1 = kb0;
}

// Final epilogue of one 32 x 32 chunk of a tile, in the COALESCED domain.  The accumulators live one row per lane; stored
// straight from there, every warp store touches 32 different lines with 16 bytes each -- partial-sector writes that cost
// ~13 us per 128 x 128 tile (tools/gemm_debug.py ksweep: 18.5 us per K = 128 tile with the stores, 4.7 us without; short-K
// products ran at 130 instead of 190 TFLOP/s).  The chunk is therefore turned through a padded shared-memory stage of the
// warp (kEpiLd), and this routine walks it by ROW SEGMENT: 8 lanes x 16 bytes = one complete 128-byte line of a row, four
// rows per step, so C is written -- and D of  C = D - mu acc  read -- in whole lines, and the column scale / triu mask /
// max|.| ride along.  A compact function of its own on purpose: the unrolled row-wise epilogue was 1.5k
// instructions per copy, and this kernel's throughput drops by 25 % when its hot code outgrows the instruction cache
// (measured twice: an IEEE division, then a three-variant epilogue, each unrolled 128 times).
// HAS_D = false: a rolled loop (compact: it is the hot path of most launches).  HAS_D = true (C = D - mu acc: Q' = Q - mu
// grad Q and the in-place updates of the triangular solves): unrolled, with the eight D rows of the chunk loaded up front
// -- taken one per iteration of a rolled loop, each waited out its own DRAM latency and those products ran at 80-90
// TFLOP/s against 170+ for the same shapes without D; unrolling the common path instead cost it 20 % (code size, and a
// 64-register save at every call), so the two are separate functions.
template <bool HAS_D>
__device__ __noinline__ float epilogue_rows(const Params& p, const float* __restrict__ stg, float* __restrict__ Cg,
                                            const float* __restrict__ Dg, const float* __restrict__ csg, int mrow0, int nbase,
                                            float mu, float oscale, int lane, float mx) {
  const int col = (lane & 7) * 4;
  const int n = nbase + col;
  const bool fast_c = nbase + 32 <= p.N && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(Cg) & 15) == 0;
  const bool fast_d = HAS_D && fast_c && (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(Dg) & 15) == 0;
  float cs[4] = {1.f, 1.f, 1.f, 1.f};
  if (csg) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n + e < p.N) {
        float sc = csg[n + e];
        if (p.colscale_sq) sc = sc * sc;
        cs[e] = p.colscale_recip ? 1.0f / sc : sc;
      }
  }
  if (p.negate) {
#pragma unroll
    for (int e = 0; e < 4; ++e) cs[e] = -cs[e];
  }
  auto one_row = [&](int i, const float4& dpre) {
    const int row = 4 * i + (lane >> 3);
    const int mm = mrow0 + row;
    if (mm >= p.M) return;
    const float4 a4 = *reinterpret_cast<const float4*>(stg + row * kEpiLd + col);
    float x[4] = {a4.x * cs[0], a4.y * cs[1], a4.z * cs[2], a4.w * cs[3]};
    if (p.triu) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (mm > n + e) x[e] = 0.f;
    }
    if (HAS_D) {
      float d[4] = {dpre.x, dpre.y, dpre.z, dpre.w};
      if (!fast_d) {
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = (n + e < p.N) ? Dg[(size_t)mm * p.ldd + n + e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = d[e] - mu * x[e];
    }
    if (p.rho_mode) {
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = x[e] * oscale;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n + e < p.N) mx = fmaxf(mx, fabsf(x[e]));
    if (fast_c) {
      *reinterpret_cast<float4*>(Cg + (size_t)mm * p.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < p.N) Cg[(size_t)mm * p.ldc + n + e] = x[e];
    }
  };
  if constexpr (HAS_D) {
    float4 dreg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = mrow0 + 4 * i + (lane >> 3);
      dreg[i] = (fast_d && mm < p.M) ? *reinterpret_cast<const float4*>(Dg + (size_t)mm * p.ldd + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) one_row(i, dreg[i]);
  } else {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) one_row(i, zero);
  }
  return mx;
}

template <int BN, bool TS>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tc_kernel(const __grid_constant__ GroupMaps maps, const __grid_constant__ Params p) {
  using C = Cfg<BN, TS>;
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::kStages * C::kStageBytes);
  uint64_t* full = bars;                         // [kStages]
  uint64_t* conv = bars + C::kStages;            // [kStages]
  uint64_t* empty = bars + 2 * C::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 3 * C::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)C::kStages * C::kStageBytes + 256);   // 4 warps x [32][kEpiLd]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // per-problem run-time flags that cancel the triangular K-range hints (every role derives the same K ranges from them)
  __shared__ int s_full[kMaxGroup];
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kMaxGroup) {
    const int g = threadIdx.x - 64;
    int f = 0;
    if (g < p.count) {
      if (p.a_full[g] && *p.a_full[g]) f |= 1;
      if (p.b_full[g] && *p.b_full[g]) f |= 2;
    }
    s_full[g] = f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 4);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_per = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_per * p.count;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * C::kStageBytes; };
  // stage layout: SS [A hi | B hi | A lo | B lo], TS [A raw | B hi | B lo]

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one_sync()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int grp = tile / tiles_per, lt = tile % tiles_per;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int m0 = tm * BM, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          const CUtensorMap* ma = prod ? &maps.a1[grp] : &maps.a0[grp];
          const CUtensorMap* mb = prod ? &maps.b1[grp] : &maps.b0[grp];
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], C::kTxBytes);
            unsigned char* sa = stage_ptr(s);
            unsigned char* sb = sa + C::kTileABytes;
            const int k0 = kb * BK;
            if (!p.a_mn[prod]) {
              tma_load_2d(sa, ma, k0, m0, &full[s]);                       // box {32 k, 128 m}
            } else {
#pragma unroll
              for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, ma, m0 + 32 * j, k0, &full[s]);   // box {32 m, 32 k}
            }
            if (!p.b_mn[prod]) {
              tma_load_2d(sb, mb, k0, n0, &full[s]);                       // box {32 k, BN n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * 4096, mb, n0 + 32 * j, k0, &full[s]);
            }
          }
        }
      }
      // producer tail: nobody waits for the stage releases of the last kStages iterations; collect them before the CTA
      // exits so that no tcgen05.commit arrival is still in flight towards shared memory that a later CTA re-initialises
      for (int j = it > C::kStages ? it - C::kStages : 0; j < it; ++j) mbar_wait(&empty[j % C::kStages], (j / C::kStages) & 1);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The whole warp walks the schedule (uniform control flow, all lanes poll the barriers); one elected lane issues.
    {
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int grp = tile / tiles_per, lt = tile % tiles_per;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int m0 = tm * BM, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        int total_kb = 0;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          total_kb += kb1 - kb0;
        }
        if (total_kb == 0) continue;                      // epilogue writes zeros without touching TMEM
        // The tensor core adds into its fp32 accumulator with truncation, a bias that grows linearly with the chain
        // length; chains are therefore cut every kChunkKB K-blocks and the chunk results are summed in registers
        // (round-to-nearest) by the epilogue warps while the next chunk accumulates in the other TMEM buffer.
        int a = 0;
        uint32_t tmem_d = 0;
        uint32_t accumulate = 0;
        int done = 0, in_chunk = 0;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          // product 1 is subtracted (a_negate); an A operand in TMEM is always K-major (the splitter transposes)
          const uint32_t idesc = make_idesc(BN, TS ? 0 : p.a_mn[prod], p.b_mn[prod], prod);
          // descriptors of K atom kk = descriptor of atom 0 + kk * step (start-address field, 16-byte units)
          const uint64_t b_step = p.b_mn[prod] ? (1024 >> 4) : (32 >> 4);
          const uint64_t a_step = p.a_mn[prod] ? (1024 >> 4) : (32 >> 4);
          for (int kb = kb0; kb < kb1; ++kb, ++it, ++done) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            if (in_chunk == 0) {
              a = tcount & 1;
              const uint32_t aph = (tcount >> 1) & 1;
              ++tcount;
              mbar_wait(&tmem_empty[a], aph ^ 1);
              tmem_d = tmem_base + (uint32_t)(a * BN);
              accumulate = 0;
            }
            mbar_wait(&conv[s], ph);
            tc_fence_after();
            const bool chunk_end = (in_chunk + 1 == kChunkKB) || (done + 1 == total_kb);
            if (elect_one_sync()) {
              const uint32_t st_base = smem_u32(stage_ptr(s));
              const uint64_t dbh0 = operand_desc(st_base + C::kBHiOff, p.b_mn[prod], 0);
              const uint64_t dbl0 = operand_desc(st_base + C::kBLoOff, p.b_mn[prod], 0);
              uint32_t acc = accumulate;
              if (p.debug & 4) {
              } else if constexpr (TS) {
                const uint32_t a_hi_t = tmem_base + (uint32_t)(C::kATmemCol0 + s * 2 * BK);   // lane 0, this stage's A hi
                const uint32_t a_lo_t = a_hi_t + BK;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                  umma_tf32_ts(tmem_d, a_lo_t + kk * UMMA_K, dbh0 + kk * b_step, idesc, acc);     // small terms first
                  umma_tf32_ts(tmem_d, a_hi_t + kk * UMMA_K, dbl0 + kk * b_step, idesc, 1u);
                  umma_tf32_ts(tmem_d, a_hi_t + kk * UMMA_K, dbh0 + kk * b_step, idesc, 1u);
                  acc = 1u;
                }
              } else {
                const uint64_t dah0 = operand_desc(st_base, p.a_mn[prod], 0);
                const uint64_t dal0 = operand_desc(st_base + C::kALoOff, p.a_mn[prod], 0);
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                  umma_tf32(tmem_d, dal0 + kk * a_step, dbh0 + kk * b_step, idesc, acc);     // small terms first
                  umma_tf32(tmem_d, dah0 + kk * a_step, dbl0 + kk * b_step, idesc, 1u);
                  umma_tf32(tmem_d, dah0 + kk * a_step, dbh0 + kk * b_step, idesc, 1u);
                  acc = 1u;
                }
              }
              umma_commit(&empty[s]);                               // frees the stage when these MMAs retire
              if (chunk_end) umma_commit(&tmem_full[a]);            // chunk complete -> epilogue
            }
            __syncwarp();
            accumulate = 1u;
            in_chunk = chunk_end ? 0 : in_chunk + 1;
          }
        }
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ===================================== 3xTF32 splitter ==================================
    const int st = threadIdx.x - kSplitWarp0 * 32;            // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int grp = tile / tiles_per, lt = tile % tiles_per;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int m0 = tm * BM, n0 = tn * BN;
      if (!tile_computed<BN>(p, m0, n0)) continue;
      for (int prod = 0; prod < 2; ++prod) {
        if (p.K[prod] <= 0) continue;
        int kb0, kb1;
        k_range<BN>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full[s], ph);
          if constexpr (TS) {
            // ---- A: this thread owns row m = TMEM lane (warp & 3) * 32 + lane; 32 K values -> registers -> TMEM
            // ---- B: 128 threads split the tile in shared memory (hi in place, lo in the second tile)
            // All shared-memory loads of the stage (A row + this thread's share of B) are issued up front so that
            // their latency is paid once per stage, not once per 16 bytes: the splitter warps run one per scheduler.
            const uint32_t sa = smem_u32(stage_ptr(s));
            const int row = (warp & 3) * 32 + lane;
            constexpr int kBVec = C::kTileBBytes / (128 * 16);        // 16-byte vectors of B per thread (8 at BN = 128)
            uint32_t araw[32];
            uint4 braw[kBVec];
            if (p.debug & 2) {
#pragma unroll
              for (int k = 0; k < 32; ++k) araw[k] = 0;
            } else if (!p.a_mn[prod]) {
              // K-major tile, SWIZZLE_128B: row m is 128 B at m*128, its 16-byte chunk c sits at position c ^ (m & 7)
              const uint32_t rp = sa + row * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 x = lds128(rp + ((c ^ (row & 7)) << 4));
                araw[4 * c] = x.x; araw[4 * c + 1] = x.y; araw[4 * c + 2] = x.z; araw[4 * c + 3] = x.w;
              }
            } else {
              // MN-major tile: four [32 k][32 m] blocks of 128-byte rows, SWIZZLE_128B_ATOM_32B (32-byte atom index
              // XOR (k & 3)); reading column m of every k row is the transposition (lanes hit 32 distinct banks)
              const uint32_t blk = sa + (row >> 5) * 4096;
              const int mb = (row & 31) * 4;
#pragma unroll
              for (int k = 0; k < 32; ++k) araw[k] = lds32(blk + k * 128 + (mb ^ ((k & 3) << 5)));
            }
            const uint32_t bh = sa + C::kBHiOff + st * 16, bl = sa + C::kBLoOff + st * 16;
            if (!(p.debug & 1)) {
#pragma unroll
              for (int j = 0; j < kBVec; ++j) braw[j] = lds128(bh + j * 2048);
            }
            {
              uint32_t ahi[32], alo[32];
#pragma unroll
              for (int k = 0; k < 32; ++k) split_tf32(araw[k], ahi[k], alo[k]);
              const uint32_t t_a = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(C::kATmemCol0 + s * 2 * BK);
              tmem_st32(t_a, ahi);
              tmem_st32(t_a + BK, alo);
            }
            if (!(p.debug & 1)) {
#pragma unroll
              for (int j = 0; j < kBVec; ++j) {
                uint4 h, l;
                split_tf32(braw[j].x, h.x, l.x); split_tf32(braw[j].y, h.y, l.y);
                split_tf32(braw[j].z, h.z, l.z); split_tf32(braw[j].w, h.w, l.w);
                sts128(bh + j * 2048, h);
                sts128(bl + j * 2048, l);
              }
            }
            tmem_st_wait();
            tc_fence_before();              // TMEM stores ordered before the MMA issuer's tcgen05.mma (via conv[s])
          } else {
            const uint32_t sa = smem_u32(stage_ptr(s));
            split_tile_smem(sa, sa + C::kTxBytes, C::kTxBytes, st);
          }
          fence_proxy_async_smem();       // generic-proxy writes -> visible to tcgen05 (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[s]);
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    int tcount = 0;
    float mx = 0.f;
    int mx_grp = -1;
    auto flush_max = [&]() {
      if (mx_grp >= 0 && p.maxabs[mx_grp]) {
        const float w = warp_max(mx);
        if (lane == 0 && w > 0.f) atomic_max_nonneg(p.maxabs[mx_grp], w);
      }
      mx = 0.f;
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int grp = tile / tiles_per, lt = tile % tiles_per;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int m0 = tm * BM, n0 = tn * BN;
      if (!tile_in_pattern<BN>(p, m0, n0)) continue;
      if (grp != mx_grp) { flush_max(); mx_grp = grp; }
      float* const Cg = p.C[grp];
      const float* Dg = p.D[grp];
      const float* const csg = p.colscale[grp];
      float mu = 0.f;
      if (Dg) mu = p.mu_max[grp] ? p.step / (*p.mu_max[grp] + p.tiny) : 1.0f;
      // rho_mode 1: Ql / rho, 2: rho * Qr (psgd.py:169-170) as ONE multiply per element -- an IEEE division here would be
      // expanded 128 times in the unrolled epilogue (+30 % SASS, measured 25 % slower end to end: instruction cache)
      const float oscale = p.rho_mode == 1 ? 1.0f / *p.rho[grp] : (p.rho_mode == 2 ? *p.rho[grp] : 1.0f);
      int total_kb = 0;
      if (!(p.triu && m0 >= n0 + BN)) {
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          total_kb += kb1 - kb0;
        }
      }
      // Q' = Q - mu (grad Q): D IS the triangular factor Q.  Below the diagonal both the product and D are zero (the
      // run-time scan vouches for Q), so those tiles -- 48 % of the launch -- are written as zeros without reading D.
      if (Dg && p.d_tri && !(s_full[grp] & 2) && total_kb == 0 && m0 >= n0 + BN) Dg = nullptr;
      if (Dg) {
        // start this tile's 64 KB of D towards L2 now; the chunk drains below take the tile's whole MMA time, so the
        // coalesced loads of the final epilogue find it there instead of waiting out DRAM once per 32-column chunk
        const int pm = m0 + q * 32 + lane;
        if (pm < p.M) {
          const float* drow = Dg + (size_t)pm * p.ldd + n0;
#pragma unroll
          for (int c = 0; c < BN / 32; ++c)
            if (n0 + c * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(drow + c * 32));
        }
      }
      float racc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) racc[j] = 0.f;
      for (int done = 0; done < total_kb; done += kChunkKB) {
        const int a = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        ++tcount;
        mbar_wait(&tmem_full[a], aph);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[a]);
      }
      // ---- final epilogue: registers -> shared-memory stage of this warp -> global (see epilogue_rows) ----------
      if (!(p.debug & 8)) {
        float* const stg = epi_stage + q * (32 * kEpiLd);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int nbase = n0 + c * 32;
          if (nbase < p.N) {                                            // uniform over the warp
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(stg + lane * kEpiLd + 4 * j4) =
                  make_float4(racc[c * 32 + 4 * j4], racc[c * 32 + 4 * j4 + 1], racc[c * 32 + 4 * j4 + 2], racc[c * 32 + 4 * j4 + 3]);
            __syncwarp();
            mx = Dg ? epilogue_rows<true>(p, stg, Cg, Dg, csg, m0 + q * 32, nbase, mu, oscale, lane, mx)
                    : epilogue_rows<false>(p, stg, Cg, Dg, csg, m0 + q * 32, nbase, mu, oscale, lane, mx);
            __syncwarp();                                               // stage free for the next chunk
          }
        }
      }
    }
    flush_max();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

// =============================================================================================
// CTA-pair variant (tcgen05 cta_group::2): one 256 x BN output tile per CLUSTER OF TWO CTAs.
//
// Why: the single-CTA kernel above is bound by shared-memory bandwidth, not by the tensor pipe (per 32-wide K stage it
// moves 144 KB through shared memory against 768 tensor-core cycles: TMA fill, the splitter's read/write of B, and the
// three MMAs each re-reading B).  With cta_group::2 the two SMs of a TPC execute ONE 256 x BN x 8 MMA: each CTA holds
// its own 128 rows of A (split in registers, stored to its own tensor memory as before) but only HALF of the B tile
// (BN/2 columns), and the hardware feeds both tensor cores from the two halves -- the B traffic per SM (TMA fill, split,
// MMA reads) halves.  Two further savings ride along:
//   * B's "hi" part is the RAW fp32 tile: kind::tf32 ignores the low 13 mantissa bits of its operands, i.e. it sees
//     trunc_tf32(b).  The splitter then only writes lo = b - trunc_tf32(b) (exact in fp32, below 2^-10 |b|); the
//     dropped term a_lo * b_lo stays below 2^-21 |a b| and has zero mean because a is split with round-to-nearest;
//   * per stage and CTA: 24 KB TMA fill + 16 KB A read + 8 KB B read + 8 KB lo write + 24 KB MMA reads = 80 KB.
// Roles and pipelines are those of gemm_tc_kernel; what changes is who signals whom:
//   full[s]        local TMA -> local splitter                       (per CTA, as before)
//   conv[s]        splitter warps of BOTH CTAs -> the LEADER's MMA warp (count 8; the peer arrives remotely)
//   empty[s]       tcgen05.commit, multicast to both CTAs -> each CTA's TMA producer
//   tmem_full[a]   tcgen05.commit, multicast                -> each CTA's epilogue warps
//   tmem_empty[a]  epilogue warps of BOTH CTAs -> the leader's MMA warp (count 8)
// Only the leader (cluster rank 0) issues MMAs; its instruction reads A from both CTAs' tensor memory and B from both
// CTAs' shared memory at the SAME addresses, and accumulates into both CTAs' tensor memory.
//
// MEASURED (profiles/r02_gemm_ablation.txt, dense 4096^3, one B200): correct in every operand layout (1.0e-6 vs float64),
// but SLOWER than the single-CTA kernel -- 1.08 ms (127 TFLOP/s fp32-equivalent) against 0.74-0.79 ms (175-187).  The
// ablation says why: with splitting AND MMAs switched off ("TMA + barriers only") the pair pipeline still needs 1.04 ms,
// the single-CTA one 0.52 ms.  The pair kernel is bound by the round trip of its stage ring, not by shared-memory
// bandwidth: every stage now crosses the cluster twice (remote mbarrier arrive of the peer's splitter, multicast commit
// back), and with A in tensor memory the ring cannot be deeper than 4 stages (2 x 128 accumulator columns + 4 x 64 A
// columns = all 512), i.e. 96 KB in flight per SM against 128 KB for the single-CTA kernel.  It therefore stays OFF by
// default (psgd_set_option("tc_pair", 1) selects it; tests/test_gpu_tensorcore.py keeps it honest); making it pay needs
// a deeper ring (A through shared memory again, or BK = 16 stages) -- left as measured, not guessed.
// =============================================================================================
constexpr int BM2 = 256;

template <int BN>
struct Cfg2 {
  static constexpr int kStages = 4;
  static constexpr int kTileABytes = BM * BK * 4;             // this CTA's 128 rows of A: 16 KB
  static constexpr int kTileBBytes = (BN / 2) * BK * 4;       // this CTA's half of the B tile: 8 KB at BN = 128
  static constexpr int kStageBytes = kTileABytes + 2 * kTileBBytes;     // [A raw | B raw (= hi) | B lo]
  static constexpr int kTxBytes = kTileABytes + kTileBBytes;
  static constexpr int kBHiOff = kTileABytes;
  static constexpr int kBLoOff = kTileABytes + kTileBBytes;
  static constexpr int kAccCols = 2 * BN;
  static constexpr int kATmemCol0 = kAccCols;
  static constexpr int kTmemCols = 512;
  static_assert(kAccCols + kStages * 2 * BK <= 512, "TMEM budget");
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                       4 * 32 * kEpiLd * 4 /*epilogue transposition stages*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// wait with cluster-scope acquire: the arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all prior tcgen05 operations of this thread retire
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// D[tmem, both CTAs] (+)= A[tmem, both CTAs: 2 x 128 rows] * B[smem: BN/2 columns from each CTA]
__device__ __forceinline__ void umma2_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc2(int n, int b_mn, int negate_a) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format  = F32
  d |= 2u << 7;                        // a_format  = TF32
  d |= 2u << 10;                       // b_format  = TF32
  d |= (uint32_t)(negate_a & 1) << 13; // a_negate
  d |= (uint32_t)(b_mn & 1) << 16;     // b_major
  d |= (uint32_t)(n >> 3) << 17;       // n_dim
  d |= (uint32_t)(BM2 >> 4) << 24;     // m_dim = 256: 128 rows in each CTA of the pair
  return d;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    gemm_tc2_kernel(const __grid_constant__ GroupMaps maps, const __grid_constant__ Params p) {
  using C = Cfg2<BN>;
  constexpr int BNH = BN / 2;                      // B columns staged by each CTA
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::kStages * C::kStageBytes);
  uint64_t* full = bars;                         // [kStages]
  uint64_t* conv = bars + C::kStages;            // [kStages]   (used in the leader)
  uint64_t* empty = bars + 2 * C::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 3 * C::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]         (used in the leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)C::kStages * C::kStageBytes + 256);   // 4 warps x [32][kEpiLd]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  __shared__ int s_full[kMaxGroup];
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kMaxGroup) {
    const int g = threadIdx.x - 64;
    int f = 0;
    if (g < p.count) {
      if (p.a_full[g] && *p.a_full[g]) f |= 1;
      if (p.b_full[g] && *p.b_full[g]) f |= 2;
    }
    s_full[g] = f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 8);                    // 4 splitter warps of each CTA
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);              // 4 epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // both CTAs' barriers are initialised before anyone signals the peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_per = p.tiles_m * p.tiles_n;   // tiles_m counts 256-row tiles here
  const int num_tiles = tiles_per * p.count;
  auto stage_ptr = [&](int s) { return smem + (size_t)s * C::kStageBytes; };

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (elect_one_sync()) {
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int grp = tile / tiles_per, lt = tile % tiles_per;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int m0 = tm * BM2, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        const int ma0 = m0 + (int)rank * BM;       // this CTA's rows of A
        const int nb0 = n0 + (int)rank * BNH;      // this CTA's columns of B
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          const CUtensorMap* ma = prod ? &maps.a1[grp] : &maps.a0[grp];
          const CUtensorMap* mb = prod ? &maps.b1[grp] : &maps.b0[grp];
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], C::kTxBytes);
            unsigned char* sa = stage_ptr(s);
            unsigned char* sb = sa + C::kBHiOff;
            const int k0 = kb * BK;
            if (!p.a_mn[prod]) {
              tma_load_2d(sa, ma, k0, ma0, &full[s]);                      // box {32 k, 128 m}
            } else {
#pragma unroll
              for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, ma, ma0 + 32 * j, k0, &full[s]);   // box {32 m, 32 k}
            }
            if (!p.b_mn[prod]) {
              tma_load_2d(sb, mb, k0, nb0, &full[s]);                      // box {32 k, BN/2 n}
            } else {
#pragma unroll
              for (int j = 0; j < BNH / 32; ++j) tma_load_2d(sb + j * 4096, mb, nb0 + 32 * j, k0, &full[s]);
            }
          }
        }
      }
      // producer tail (BOTH CTAs): the leader's multicast tcgen05.commit arrivals on empty[] of the last kStages
      // iterations are asynchronous and waited for by nobody; without this a late arrival can land after this CTA has
      // exited, on the freshly initialised barrier of the next launch's CTA (same shared-memory layout), which then
      // refills a stage the tensor core is still reading -- seen as rare, timing-dependent corruption of back-to-back
      // launches (the triangular-solve recursion)
      for (int j = it > C::kStages ? it - C::kStages : 0; j < it; ++j) mbar_wait(&empty[j % C::kStages], (j / C::kStages) & 1);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (rank == 0) {
      int it = 0, tcount = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int grp = tile / tiles_per, lt = tile % tiles_per;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int m0 = tm * BM2, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        int total_kb = 0;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          total_kb += kb1 - kb0;
        }
        if (total_kb == 0) continue;
        int a = 0;
        uint32_t tmem_d = 0;
        uint32_t accumulate = 0;
        int done = 0, in_chunk = 0;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          const uint32_t idesc = make_idesc2(BN, p.b_mn[prod], prod);
          const uint64_t b_step = p.b_mn[prod] ? (1024 >> 4) : (32 >> 4);
          for (int kb = kb0; kb < kb1; ++kb, ++it, ++done) {
            const int s = it % C::kStages;
            const uint32_t ph = (it / C::kStages) & 1;
            if (in_chunk == 0) {
              a = tcount & 1;
              const uint32_t aph = (tcount >> 1) & 1;
              ++tcount;
              mbar_wait_cluster(&tmem_empty[a], aph ^ 1);
              tmem_d = tmem_base + (uint32_t)(a * BN);
              accumulate = 0;
            }
            mbar_wait_cluster(&conv[s], ph);
            tc_fence_after();
            const bool chunk_end = (in_chunk + 1 == kChunkKB) || (done + 1 == total_kb);
            if (elect_one_sync()) {
              const uint32_t st_base = smem_u32(stage_ptr(s));
              const uint64_t dbh0 = operand_desc(st_base + C::kBHiOff, p.b_mn[prod], 0);
              const uint64_t dbl0 = operand_desc(st_base + C::kBLoOff, p.b_mn[prod], 0);
              uint32_t acc = accumulate;
              if (!(p.debug & 4)) {
                const uint32_t a_hi_t = tmem_base + (uint32_t)(C::kATmemCol0 + s * 2 * BK);
                const uint32_t a_lo_t = a_hi_t + BK;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                  umma2_tf32_ts(tmem_d, a_lo_t + kk * UMMA_K, dbh0 + kk * b_step, idesc, acc);     // small terms first
                  umma2_tf32_ts(tmem_d, a_hi_t + kk * UMMA_K, dbl0 + kk * b_step, idesc, 1u);
                  umma2_tf32_ts(tmem_d, a_hi_t + kk * UMMA_K, dbh0 + kk * b_step, idesc, 1u);
                  acc = 1u;
                }
              }
              umma_commit2(&empty[s]);                              // frees the stage in both CTAs
              if (chunk_end) umma_commit2(&tmem_full[a]);           // chunk complete -> both CTAs' epilogue warps
            }
            __syncwarp();
            accumulate = 1u;
            in_chunk = chunk_end ? 0 : in_chunk + 1;
          }
        }
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ===================================== 3xTF32 splitter (both CTAs) =======================
    const int st = threadIdx.x - kSplitWarp0 * 32;            // 0..127
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int grp = tile / tiles_per, lt = tile % tiles_per;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int m0 = tm * BM2, n0 = tn * BN;
      if (!tile_computed<BN>(p, m0, n0)) continue;
      for (int prod = 0; prod < 2; ++prod) {
        if (p.K[prod] <= 0) continue;
        int kb0, kb1;
        k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full[s], ph);
          const uint32_t sa = smem_u32(stage_ptr(s));
          const int row = (warp & 3) * 32 + lane;
          constexpr int kBVec = C::kTileBBytes / (128 * 16);        // 16-byte vectors of B per thread (4 at BN = 128)
          uint32_t araw[32];
          uint4 braw[kBVec];
          if (p.debug & 2) {
#pragma unroll
            for (int k = 0; k < 32; ++k) araw[k] = 0;
          } else if (!p.a_mn[prod]) {
            const uint32_t rp = sa + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 x = lds128(rp + ((c ^ (row & 7)) << 4));
              araw[4 * c] = x.x; araw[4 * c + 1] = x.y; araw[4 * c + 2] = x.z; araw[4 * c + 3] = x.w;
            }
          } else {
            const uint32_t blk = sa + (row >> 5) * 4096;
            const int mb = (row & 31) * 4;
#pragma unroll
            for (int k = 0; k < 32; ++k) araw[k] = lds32(blk + k * 128 + (mb ^ ((k & 3) << 5)));
          }
          const uint32_t bh = sa + C::kBHiOff + st * 16, bl = sa + C::kBLoOff + st * 16;
          if (!(p.debug & 1)) {
#pragma unroll
            for (int j = 0; j < kBVec; ++j) braw[j] = lds128(bh + j * 2048);
          }
          {
            uint32_t ahi[32], alo[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) split_tf32(araw[k], ahi[k], alo[k]);
            const uint32_t t_a = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(C::kATmemCol0 + s * 2 * BK);
            tmem_st32(t_a, ahi);
            tmem_st32(t_a + BK, alo);
          }
          if (!(p.debug & 1)) {
            // lo = b - trunc_tf32(b): the raw tile itself serves as "hi" (the tensor core drops the low 13 bits)
#pragma unroll
            for (int j = 0; j < kBVec; ++j) {
              uint4 l;
              l.x = __float_as_uint(__uint_as_float(braw[j].x) - __uint_as_float(braw[j].x & 0xffffe000u));
              l.y = __float_as_uint(__uint_as_float(braw[j].y) - __uint_as_float(braw[j].y & 0xffffe000u));
              l.z = __float_as_uint(__uint_as_float(braw[j].z) - __uint_as_float(braw[j].z & 0xffffe000u));
              l.w = __float_as_uint(__uint_as_float(braw[j].w) - __uint_as_float(braw[j].w & 0xffffe000u));
              sts128(bl + j * 2048, l);
            }
          }
          tmem_st_wait();
          tc_fence_before();              // TMEM stores ordered before the leader's tcgen05.mma (via conv[s])
          fence_proxy_async_smem();       // generic-proxy writes of lo -> visible to tcgen05 (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&conv[s]);
            else mbar_arrive_remote(&conv[s], 0);
          }
        }
      }
    }
  } else {
    // ===================================== epilogue (both CTAs) ==============================
    const int q = warp & 3;
    int tcount = 0;
    float mx = 0.f;
    int mx_grp = -1;
    auto flush_max = [&]() {
      if (mx_grp >= 0 && p.maxabs[mx_grp]) {
        const float w = warp_max(mx);
        if (lane == 0 && w > 0.f) atomic_max_nonneg(p.maxabs[mx_grp], w);
      }
      mx = 0.f;
    };
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int grp = tile / tiles_per, lt = tile % tiles_per;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int m0 = tm * BM2, n0 = tn * BN;
      if (!tile_in_pattern<BN>(p, m0, n0)) continue;
      if (grp != mx_grp) { flush_max(); mx_grp = grp; }
      float* const Cg = p.C[grp];
      const float* const Dg = p.D[grp];
      const float* const csg = p.colscale[grp];
      float mu = 0.f;
      if (Dg) mu = p.mu_max[grp] ? p.step / (*p.mu_max[grp] + p.tiny) : 1.0f;
      // rho_mode 1: Ql / rho, 2: rho * Qr (psgd.py:169-170) as ONE multiply per element -- an IEEE division here would be
      // expanded 128 times in the unrolled epilogue (+30 % SASS, measured 25 % slower end to end: instruction cache)
      const float oscale = p.rho_mode == 1 ? 1.0f / *p.rho[grp] : (p.rho_mode == 2 ? *p.rho[grp] : 1.0f);
      int total_kb = 0;
      if (!(p.triu && m0 >= n0 + BN)) {
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1, s_full[grp]);
          total_kb += kb1 - kb0;
        }
      }
      float racc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) racc[j] = 0.f;
      for (int done = 0; done < total_kb; done += kChunkKB) {
        const int a = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        ++tcount;
        mbar_wait(&tmem_full[a], aph);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&tmem_empty[a]);
          else mbar_arrive_remote(&tmem_empty[a], 0);
        }
      }
      if (!(p.debug & 8)) {
        float* const stg = epi_stage + q * (32 * kEpiLd);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int nbase = n0 + c * 32;
          if (nbase < p.N) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(stg + lane * kEpiLd + 4 * j4) =
                  make_float4(racc[c * 32 + 4 * j4], racc[c * 32 + 4 * j4 + 1], racc[c * 32 + 4 * j4 + 2], racc[c * 32 + 4 * j4 + 3]);
            __syncwarp();
            mx = Dg ? epilogue_rows<true>(p, stg, Cg, Dg, csg, m0 + (int)rank * BM + q * 32, nbase, mu, oscale, lane, mx)
                    : epilogue_rows<false>(p, stg, Cg, Dg, csg, m0 + (int)rank * BM + q * 32, nbase, mu, oscale, lane, mx);
            __syncwarp();
          }
        }
      }
    }
    flush_max();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // the peer's smem / tensor memory / barriers stay valid until both CTAs are done
  if (warp == 1) tmem_dealloc2(tmem_base, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Row-major [rows, cols] fp32 matrix with leading dimension ld; box = {32 floats along cols, box_rows rows}.
static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, bool mn_major) {
  EncodeTiledFn fn = get_encode_fn();
  PSGD_REQUIRE(fn != nullptr, PSGD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PSGD_REQUIRE(r == CUDA_SUCCESS, PSGD_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows=%d cols=%d ld=%d)", (int)r,
               rows, cols, ld);
  return PSGD_OK;
}

static bool operand_ok(const float* p, int ld) { return p && aligned16(p) && (ld % 4) == 0; }

bool gemm_tc_supported(const la::Gemm& g) {
  if (g.M < 1 || g.N < 1 || g.K < 1) return false;
  if (!operand_ok(g.A, g.lda) || !operand_ok(g.B, g.ldb)) return false;
  if (g.K2 > 0 && (!operand_ok(g.A2, g.lda2) || !operand_ok(g.B2, g.ldb2))) return false;
  return true;
}

static bool same_shape(const la::Gemm& x, const la::Gemm& y) {
  return x.M == y.M && x.N == y.N && x.K == y.K && x.K2 == y.K2 && x.lda == y.lda && x.ldb == y.ldb && x.lda2 == y.lda2 &&
         x.ldb2 == y.ldb2 && x.ldc == y.ldc && x.ldd == y.ldd && x.ta == y.ta && x.tb == y.tb && x.ta2 == y.ta2 &&
         x.tb2 == y.tb2 && x.triu == y.triu && x.a_tri == y.a_tri && x.b_tri == y.b_tri && x.step == y.step &&
         x.tiny == y.tiny && x.colscale_recip == y.colscale_recip && x.colscale_sq == y.colscale_sq &&
         (x.D != nullptr) == (y.D != nullptr) && (x.K2 > 0) == (y.K2 > 0) && x.pair_b == y.pair_b &&
         x.pair_kind == y.pair_kind && x.negate == y.negate && x.rho_mode == y.rho_mode &&
         (x.a_full != nullptr) == (y.a_full != nullptr) && (x.b_full != nullptr) == (y.b_full != nullptr) &&
         x.d_tri == y.d_tri;
}

// gs[0..count): identical shapes/flags, count <= kMaxGroup
template <int BN, bool TS>
static int launch_impl(psgd_ctx* ctx, const la::Gemm* gs, int count) {
  using C = Cfg<BN, TS>;
  const la::Gemm& g = gs[0];
  static thread_local Params p;            // large by-value kernel parameters: keep them off the stack
  static thread_local GroupMaps maps;
  p = Params{};
  p.M = g.M; p.N = g.N;
  p.K[0] = g.K; p.K[1] = g.K2;
  // op(A) is [M,K]: !ta -> A stored [M,K] (K-major); ta -> A stored [K,M] (MN-major)
  p.a_mn[0] = g.ta ? 1 : 0;
  // op(B) is [K,N]: tb -> B stored [N,K] (K-major); !tb -> B stored [K,N] (MN-major)
  p.b_mn[0] = g.tb ? 0 : 1;
  p.a_mn[1] = g.ta2 ? 1 : 0;
  p.b_mn[1] = g.tb2 ? 0 : 1;
  p.a_tri = g.a_tri; p.b_tri = g.b_tri;
  p.ldc = g.ldc; p.ldd = g.ldd; p.triu = g.triu ? 1 : 0;
  p.step = g.step; p.tiny = g.tiny;
  p.colscale_recip = g.colscale_recip; p.colscale_sq = g.colscale_sq;
  p.tiles_m = (g.M + BM - 1) / BM;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.count = count;
  p.debug = ctx->opt_tc_debug;
  p.epi = ctx->opt_tc_epi;
  p.d_tri = (g.d_tri && g.b_full && g.b_tri == 1 && g.a_tri == 1) ? 1 : 0;
  p.pair_b = g.pair_b; p.pair_kind = g.pair_kind; p.negate = g.negate ? 1 : 0;
  p.rho_mode = g.rho ? g.rho_mode : 0;
  double work = 0.0;
  for (int i = 0; i < count; ++i) {
    const la::Gemm& q = gs[i];
    p.C[i] = q.C; p.maxabs[i] = q.maxabs; p.D[i] = q.D; p.mu_max[i] = q.mu_max; p.colscale[i] = q.colscale;
    p.a_full[i] = q.a_full; p.b_full[i] = q.b_full; p.rho[i] = q.rho;
    work += 2.0 * q.M * q.N * ((double)q.K + q.K2);
    for (int prod = 0; prod < 2; ++prod) {
      const float* A = prod ? q.A2 : q.A;
      const float* B = prod ? q.B2 : q.B;
      const int lda = prod ? q.lda2 : q.lda, ldb = prod ? q.ldb2 : q.ldb;
      const int K = prod ? q.K2 : q.K;
      CUtensorMap* ta = prod ? &maps.a1[i] : &maps.a0[i];
      CUtensorMap* tb = prod ? &maps.b1[i] : &maps.b0[i];
      if (K <= 0) continue;
      if (!p.a_mn[prod]) PSGD_RETURN_IF(make_map(ta, A, q.M, K, lda, BM, false));     // [M,K], box {32k, 128m}
      else               PSGD_RETURN_IF(make_map(ta, A, K, q.M, lda, BK, true));      // [K,M], box {32m, 32k}
      if (!p.b_mn[prod]) PSGD_RETURN_IF(make_map(tb, B, q.N, K, ldb, BN, false));     // [N,K], box {32k, BN n}
      else               PSGD_RETURN_IF(make_map(tb, B, K, q.N, ldb, BK, true));      // [K,N], box {32n, 32k}
    }
  }
  auto kern = gemm_tc_kernel<BN, TS>;
  static DeviceOnce attr_done;             // one per template instantiation
  if (!attr_done.done(ctx->device)) {
    PSGD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    attr_done.set(ctx->device);
  }
  int grid = p.tiles_m * p.tiles_n * count;
  if (grid > ctx->num_sms) grid = ctx->num_sms;
  if (ctx->opt_profile == 2) {
    // psgd_set_option("profile", 2): report the flops the launch EXECUTES (K blocks left after the triangular clipping,
    // tiles left after triu / block-pair skipping; one fp32-equivalent multiply-add per 3xTF32 triple) instead of the
    // dense count of the op it implements
    double kbs = 0.0;
    for (int tm = 0; tm < p.tiles_m; ++tm)
      for (int tn = 0; tn < p.tiles_n; ++tn) {
        const int m0 = tm * BM, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN>(p, prod, m0, n0, kb0, kb1);
          kbs += kb1 - kb0;
        }
      }
    work = kbs * 2.0 * BM * BN * BK * count;
  }
  ProfScope prof(ctx, PSGD_K_GEMM, work);
  kern<<<grid, kThreads, C::kSmemBytes, ctx->stream>>>(maps, p);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// CTA-pair launch (gemm_tc2_kernel): same problem description, 256-row tiles, clusters of two CTAs.
template <int BN>
static int launch_pair(psgd_ctx* ctx, const la::Gemm* gs, int count) {
  using C = Cfg2<BN>;
  const la::Gemm& g = gs[0];
  static thread_local Params p;
  static thread_local GroupMaps maps;
  p = Params{};
  p.M = g.M; p.N = g.N;
  p.K[0] = g.K; p.K[1] = g.K2;
  p.a_mn[0] = g.ta ? 1 : 0;
  p.b_mn[0] = g.tb ? 0 : 1;
  p.a_mn[1] = g.ta2 ? 1 : 0;
  p.b_mn[1] = g.tb2 ? 0 : 1;
  p.a_tri = g.a_tri; p.b_tri = g.b_tri;
  p.ldc = g.ldc; p.ldd = g.ldd; p.triu = g.triu ? 1 : 0;
  p.step = g.step; p.tiny = g.tiny;
  p.colscale_recip = g.colscale_recip; p.colscale_sq = g.colscale_sq;
  p.tiles_m = (g.M + BM2 - 1) / BM2;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.count = count;
  p.debug = ctx->opt_tc_debug;
  p.negate = g.negate ? 1 : 0;
  p.rho_mode = g.rho ? g.rho_mode : 0;
  double work = 0.0;
  for (int i = 0; i < count; ++i) {
    const la::Gemm& q = gs[i];
    p.C[i] = q.C; p.maxabs[i] = q.maxabs; p.D[i] = q.D; p.mu_max[i] = q.mu_max; p.colscale[i] = q.colscale;
    p.a_full[i] = q.a_full; p.b_full[i] = q.b_full; p.rho[i] = q.rho;
    work += 2.0 * q.M * q.N * ((double)q.K + q.K2);
    for (int prod = 0; prod < 2; ++prod) {
      const float* A = prod ? q.A2 : q.A;
      const float* B = prod ? q.B2 : q.B;
      const int lda = prod ? q.lda2 : q.lda, ldb = prod ? q.ldb2 : q.ldb;
      const int K = prod ? q.K2 : q.K;
      CUtensorMap* ta = prod ? &maps.a1[i] : &maps.a0[i];
      CUtensorMap* tb = prod ? &maps.b1[i] : &maps.b0[i];
      if (K <= 0) continue;
      if (!p.a_mn[prod]) PSGD_RETURN_IF(make_map(ta, A, q.M, K, lda, BM, false));          // [M,K], box {32k, 128m}
      else               PSGD_RETURN_IF(make_map(ta, A, K, q.M, lda, BK, true));           // [K,M], box {32m, 32k}
      if (!p.b_mn[prod]) PSGD_RETURN_IF(make_map(tb, B, q.N, K, ldb, BN / 2, false));      // [N,K], box {32k, BN/2 n}
      else               PSGD_RETURN_IF(make_map(tb, B, K, q.N, ldb, BK, true));           // [K,N], box {32n, 32k}
    }
  }
  auto kern = gemm_tc2_kernel<BN>;
  static DeviceOnce attr_done;
  if (!attr_done.done(ctx->device)) {
    PSGD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
    attr_done.set(ctx->device);
  }
  // persistent clusters: never launch more clusters than can be co-resident (a second wave would double the time of a
  // statically scheduled tile loop).  Not every TPC has both SMs enabled, so this can be fewer than num_sms / 2.
  static int max_clusters[64] = {0};
  const int dv = ctx->device & 63;
  if (max_clusters[dv] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctx->num_sms / 2 * 2);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    if (e != cudaSuccess || nc <= 0) { (void)cudaGetLastError(); nc = ctx->num_sms / 2; }
    if (nc > ctx->num_sms / 2) nc = ctx->num_sms / 2;
    max_clusters[dv] = nc;
  }
  int pairs = p.tiles_m * p.tiles_n * count;
  if (pairs > max_clusters[dv]) pairs = max_clusters[dv];
  if (ctx->opt_profile == 2) {
    // executed flops, counted per CTA tile (128 rows): the K blocks of the PAIR tile are executed by both halves
    double kbs = 0.0;
    for (int tm = 0; tm < p.tiles_m; ++tm)
      for (int tn = 0; tn < p.tiles_n; ++tn) {
        const int m0 = tm * BM2, n0 = tn * BN;
        if (!tile_computed<BN>(p, m0, n0)) continue;
        for (int prod = 0; prod < 2; ++prod) {
          if (p.K[prod] <= 0) continue;
          int kb0, kb1;
          k_range<BN, BM2>(p, prod, m0, n0, kb0, kb1);
          kbs += kb1 - kb0;
        }
      }
    work = kbs * 2.0 * BM2 * BN * BK * count;
  }
  ProfScope prof(ctx, PSGD_K_GEMM, work);
  kern<<<2 * pairs, kThreads, C::kSmemBytes, ctx->stream>>>(maps, p);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// opt_tc_mode: 1 (default) = A operand through tensor memory ("TS"), 0 = both operands from shared memory ("SS").
// opt_tc_pair: 1 = CTA-pair kernel (cta_group::2, 256-row tiles) whenever the problem has at least 256 rows and is not a
// block-pair (triangular inverse doubling) launch; 0 (default) = single-CTA kernel only (see the measurements above).
template <int BN>
static int launch(psgd_ctx* ctx, const la::Gemm* gs, int count) {
  const int seq = ctx->tc_launch_seq++;
  const bool sel = ctx->opt_tc_pair_sel < 0 || ctx->opt_tc_pair_sel == seq;
  if (ctx->opt_tc_pair_sel >= 0 && sel && getenv("PSGD_TC_TRACE"))
    fprintf(stderr, "[tc %d] M=%d N=%d K=%d K2=%d ta=%d tb=%d triu=%d a_tri=%d b_tri=%d D=%d maxabs=%d rho=%d cs=%d pair_b=%d cnt=%d\n", seq,
            gs[0].M, gs[0].N, gs[0].K, gs[0].K2, (int)gs[0].ta, (int)gs[0].tb, (int)gs[0].triu, gs[0].a_tri, gs[0].b_tri,
            gs[0].D != nullptr, gs[0].maxabs != nullptr, gs[0].rho_mode, gs[0].colscale != nullptr, gs[0].pair_b, count);
  if (ctx->opt_tc_pair && sel && ctx->opt_tc_mode && gs[0].pair_b == 0 && gs[0].M >= BM2) return launch_pair<BN>(ctx, gs, count);
  return ctx->opt_tc_mode ? launch_impl<BN, true>(ctx, gs, count) : launch_impl<BN, false>(ctx, gs, count);
}

// ---------------------------------------------------------------------------------------------
// split-K: a product with a small output and a long contraction (the 256 x 256 Gram differences of the NMT embedding
// layers have K = 9414 / 4935: three 128 x 128 tiles, i.e. three SMs, walking 590 K blocks each).  The K range is cut
// into S parts that run as S grouped problems of ONE launch into S partial outputs; a streaming kernel then sums the
// parts in fixed order (deterministic) and applies the epilogue.  Only for problems without triangular K-range hints.
// ---------------------------------------------------------------------------------------------
struct FinishArgs {
  const float* part;        // [nparts][M * N]
  int nparts, nneg;         // the LAST nneg parts are subtracted (second product)
  int M, N, ldc, ldd;
  float* C;
  const float* D;
  const float* mu_max;
  float step, tiny;
  float* maxabs;
  const float* colscale;
  int colscale_recip, colscale_sq, triu, negate;
  const float* rho;
  int rho_mode;
};

__global__ void __launch_bounds__(256) splitk_finish_kernel(const FinishArgs a) {
  const size_t MN = (size_t)a.M * a.N;
  float mu = 0.f;
  if (a.D) mu = a.mu_max ? a.step / (*a.mu_max + a.tiny) : 1.0f;
  const float oscale = a.rho_mode == 1 ? 1.0f / *a.rho : (a.rho_mode == 2 ? *a.rho : 1.0f);
  float mx = 0.f;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < MN; e += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(e / a.N), n = (int)(e % a.N);
    float x = 0.f;
    for (int s = 0; s < a.nparts - a.nneg; ++s) x += a.part[(size_t)s * MN + e];
    float y = 0.f;
    for (int s = a.nparts - a.nneg; s < a.nparts; ++s) y += a.part[(size_t)s * MN + e];
    x = x - y;
    if (a.negate) x = -x;
    if (a.colscale) {
      float sc = a.colscale[n];
      if (a.colscale_sq) sc = sc * sc;
      x = a.colscale_recip ? x * (1.0f / sc) : x * sc;
    }
    if (a.triu && m > n) x = 0.f;
    if (a.D) x = a.D[(size_t)m * a.ldd + n] - mu * x;
    if (a.rho_mode) x = x * oscale;
    mx = fmaxf(mx, fabsf(x));
    a.C[(size_t)m * a.ldc + n] = x;
  }
  if (a.maxabs) {
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_nonneg(a.maxabs, mx);
  }
}

// Number of parts for problem g run `count` times in one group (0 = no split).
static int splitk_parts(const psgd_ctx* ctx, const la::Gemm& g, int count) {
  if (!ctx->opt_tc_splitk || g.a_tri || g.b_tri || g.pair_b || g.a_full || g.b_full) return 0;
  const int tiles = ((g.M + BM - 1) / BM) * ((g.N + 127) / 128) * count;
  const int kmax = g.K > g.K2 ? g.K : g.K2;
  if (tiles * 3 > ctx->num_sms || kmax < 2048) return 0;
  const int nprod = g.K2 > 0 ? 2 : 1;
  int S = ctx->num_sms / tiles;
  if (S > kmax / 512) S = kmax / 512;
  if (S > kMaxGroup / (count * nprod)) S = kMaxGroup / (count * nprod);
  return S >= 2 ? S : 0;
}

template <int BN>
static int launch(psgd_ctx* ctx, const la::Gemm* gs, int count);

static int gemm_splitk(psgd_ctx* ctx, const la::Gemm* gs, int count, int S) {
  const la::Gemm& g0 = gs[0];
  const size_t MN = (size_t)g0.M * g0.N;
  const int nprod = g0.K2 > 0 ? 2 : 1;
  // parts per problem: S for each product (a part is dropped when its K range is empty)
  std::vector<la::Gemm> subs;
  std::vector<int> nparts(count, 0), nneg(count, 0);
  const int slot = ctx->stream_slot();
  float* scratch = nullptr;
  PSGD_RETURN_IF(ctx->reserve_aux(slot, (size_t)count * S * nprod * MN * sizeof(float), &scratch));
  for (int prod = 0; prod < nprod; ++prod) {
    const int K = prod ? g0.K2 : g0.K;
    const int Kc = ((K + S - 1) / S + BK - 1) / BK * BK;
    // full-size parts first (one grouped launch), the remainder part (shorter K) as a launch of its own
    for (int pass = 0; pass < 2; ++pass) {
      subs.clear();
      for (int i = 0; i < count; ++i) {
        const la::Gemm& g = gs[i];
        const float* A = prod ? g.A2 : g.A;
        const float* B = prod ? g.B2 : g.B;
        const int lda = prod ? g.lda2 : g.lda, ldb = prod ? g.ldb2 : g.ldb;
        const bool ta = prod ? g.ta2 : g.ta, tb = prod ? g.tb2 : g.tb;
        for (int s = 0; s < S; ++s) {
          const int k0 = s * Kc;
          if (k0 >= K) break;
          const int Ks = K - k0 < Kc ? K - k0 : Kc;
          if ((Ks == Kc) != (pass == 0)) continue;
          la::Gemm q;
          q.M = g.M; q.N = g.N; q.K = Ks;
          q.A = ta ? A + (size_t)k0 * lda : A + k0; q.lda = lda; q.ta = ta;
          q.B = tb ? B + k0 : B + (size_t)k0 * ldb; q.ldb = ldb; q.tb = tb;
          // part index within problem i: product 0 parts first, then product 1 parts (subtracted by the finish kernel)
          const int pidx = nparts[i]++;
          if (prod) nneg[i]++;
          q.C = scratch + ((size_t)i * S * nprod + pidx) * MN; q.ldc = g.N;
          subs.push_back(q);
        }
      }
      for (size_t b = 0; b < subs.size(); b += kMaxGroup) {
        const int cnt = (int)std::min<size_t>(kMaxGroup, subs.size() - b);
        PSGD_RETURN_IF(launch<128>(ctx, subs.data() + b, cnt));
      }
    }
  }
  for (int i = 0; i < count; ++i) {
    const la::Gemm& g = gs[i];
    FinishArgs a{};
    a.part = scratch + (size_t)i * S * nprod * MN; a.nparts = nparts[i]; a.nneg = nneg[i];
    a.M = g.M; a.N = g.N; a.ldc = g.ldc; a.ldd = g.ldd; a.C = g.C; a.D = g.D; a.mu_max = g.mu_max; a.step = g.step; a.tiny = g.tiny;
    a.maxabs = g.maxabs; a.colscale = g.colscale; a.colscale_recip = g.colscale_recip; a.colscale_sq = g.colscale_sq;
    a.triu = g.triu ? 1 : 0; a.negate = g.negate ? 1 : 0; a.rho = g.rho; a.rho_mode = g.rho ? g.rho_mode : 0;
    int blocks = (int)((MN + 255) / 256);
    if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
    splitk_finish_kernel<<<blocks, 256, 0, ctx->stream>>>(a);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

int gemm_tc(psgd_ctx* ctx, const la::Gemm& g) {
  PSGD_REQUIRE(gemm_tc_supported(g), PSGD_ERR_BAD_SHAPE,
               "tcgen05 GEMM needs 16-byte aligned operands with leading dimensions that are multiples of 4");
  if (g.M <= 0 || g.N <= 0) return PSGD_OK;
  return gemm_many(ctx, &g, 1, true);
}

static bool want_tc(const psgd_ctx* ctx, const la::Gemm& g) {
  const bool big = g.M >= 256 && g.N >= 256 && g.K >= 256;
  return gemm_tc_supported(g) && (ctx->opt_gemm_path == 2 || (ctx->opt_gemm_path == 0 && big));
}

// Same GEMM for many layers.  Problems that are tensor-core eligible and share shape/flags with their neighbours go out
// as grouped launches; everything else runs one by one.
int gemm_many(psgd_ctx* ctx, const la::Gemm* gs, int count, bool force_tc) {
  int i = 0;
  while (i < count) {
    const bool tcok = gemm_tc_supported(gs[i]) && (force_tc || want_tc(ctx, gs[i]));
    if (!tcok) {
      PSGD_RETURN_IF(gemm_auto(ctx, gs[i]));
      ++i;
      continue;
    }
    int j = i + 1;
    while (j < count && j - i < kMaxGroup && same_shape(gs[i], gs[j]) && gemm_tc_supported(gs[j])) ++j;
    if (gs[i].M > 0 && gs[i].N > 0) {
      const int S = splitk_parts(ctx, gs[i], j - i);
      if (S) PSGD_RETURN_IF(gemm_splitk(ctx, gs + i, j - i, S));
      else PSGD_RETURN_IF(launch<128>(ctx, gs + i, j - i));
    }
    i = j;
  }
  return PSGD_OK;
}

int gemm_auto(psgd_ctx* ctx, const la::Gemm& g) {
  if (want_tc(ctx, g)) return gemm_tc(ctx, g);
  ProfScope prof(ctx, PSGD_K_GEMM_SIMT, 2.0 * g.M * g.N * ((double)g.K + g.K2));
  return la::gemm_simt(ctx, g);
}

// ---------------------------------------------------------------------------------------------
// triangular solves: recursive blocking, grouped over layers.  Replaces tf.linalg.triangular_solve (psgd.py:174,
// :233, :298) for large factors.  The two half-size solves recurse down to kTrsmBase-wide diagonal blocks; a diagonal
// block is applied as a GEMM with its explicit inverse (computed once per solve by a small SIMT kernel, the approach
// of blocked BLAS TRSMs), and everything off the diagonal is one large tcgen05 GEMM per recursion level, so every
// multiply-add of the solve runs on the tensor cores and every step is ONE launch for the whole group of layers.
// ---------------------------------------------------------------------------------------------
constexpr int kInvBlock = 128;         // diagonal blocks inverted by back-substitution in shared memory

static int round_up(int x, int m) { return (x + m - 1) / m * m; }
// per problem: Z (n x n, explicit inverses of the diagonal base blocks) + T (n x n, scratch of the doubling steps)
size_t trsm_scratch_floats(int n) { return 2 * (size_t)round_up(n, kInvBlock) * round_up(n, kInvBlock); }

// Z[b0:b0+jb, b0:b0+jb] = inv(Q[b0:b0+jb, b0:b0+jb]) for every 128-wide diagonal block b (only the upper triangle of Q is
// read; the strictly lower part of the block is written as zeros); grid = blocks.
constexpr int kInvBatch = 64;           // problems per launch of tri_inv_blocks_kernel (pointers travel as kernel parameters)
struct InvBatch {
  const float* Q[kInvBatch];
  float* Z[kInvBatch];
};

// grid = (blocks, problems): all layers of a group in ONE launch (one launch per layer left 32 CTAs on 148 SMs and cost
// 24 x 97 us per solve of the 24 x 4096^2 stack)
__global__ void __launch_bounds__(kInvBlock) tri_inv_blocks_kernel(const __grid_constant__ InvBatch batch, int ldq, int n,
                                                                   int ldz) {
  extern __shared__ float sm[];
  float (*T)[kInvBlock + 1] = reinterpret_cast<float (*)[kInvBlock + 1]>(sm);
  float (*Zs)[kInvBlock + 1] = reinterpret_cast<float (*)[kInvBlock + 1]>(sm + kInvBlock * (kInvBlock + 1));
  const float* __restrict__ Q = batch.Q[blockIdx.y];
  float* __restrict__ Z = batch.Z[blockIdx.y];
  const int b0 = blockIdx.x * kInvBlock;
  const int jb = min(kInvBlock, n - b0);
  const int j = threadIdx.x;
  for (int i = 0; i < kInvBlock; ++i) {
    T[i][j] = (i < jb && j < jb && i <= j) ? Q[(size_t)(b0 + i) * ldq + b0 + j] : (i == j ? 1.f : 0.f);
    Zs[i][j] = 0.f;
  }
  __syncthreads();
  if (j < jb) {
    // column j of the inverse by back substitution, rows i = j .. 0 (each thread only reads its own column of Zs)
    Zs[j][j] = 1.0f / T[j][j];
    for (int i = j - 1; i >= 0; --i) {
      float s = 0.f;
      for (int k = i + 1; k <= j; ++k) s = fmaf(T[i][k], Zs[k][j], s);
      Zs[i][j] = -s / T[i][i];
    }
  }
  __syncthreads();
  if (j < jb)
    for (int i = 0; i < jb; ++i) Z[(size_t)(b0 + i) * ldz + b0 + j] = Zs[i][j];
  // The block to the LEFT of every odd diagonal block is structurally zero (Z is upper triangular) but lives in
  // uninitialised scratch.  128-row tiles never read it under the triangular K-range hints; the CTA-pair kernel's
  // 256-row tiles do (one K range serves both halves of the tile: rows m0 .. m0+127 also see k up to m0+255), so it is
  // written here -- found as stale-workspace garbage in the left leaves of the solve once a 128-wide run had gone before.
  if ((blockIdx.x & 1) && j < kInvBlock)
    for (int i = 0; i < jb; ++i) Z[(size_t)(b0 + i) * ldz + (b0 - kInvBlock) + j] = 0.f;
}

static int invert_diag_blocks(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int n, int ldz) {
  const int blocks = (n + kInvBlock - 1) / kInvBlock;
  const size_t smem = 2 * (size_t)kInvBlock * (kInvBlock + 1) * sizeof(float);
  static DeviceOnce attr_done;
  if (!attr_done.done(ctx->device)) {
    PSGD_CUDA_CHECK(cudaFuncSetAttribute(tri_inv_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.set(ctx->device);
  }
  for (int t0 = 0; t0 < count; t0 += kInvBatch) {
    const int cnt = count - t0 < kInvBatch ? count - t0 : kInvBatch;
    InvBatch batch{};
    for (int t = 0; t < cnt; ++t) { batch.Q[t] = ts[t0 + t].Q; batch.Z[t] = ts[t0 + t].zinv; }
    ProfScope prof(ctx, PSGD_K_TRSM, (double)cnt * n * kInvBlock * kInvBlock / 3.0);
    tri_inv_blocks_kernel<<<dim3(blocks, cnt), kInvBlock, smem, ctx->stream>>>(batch, ldq, n, ldz);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

// Base-block size of the solves: the recursion stops at `base`-wide diagonal blocks, which are applied as ONE GEMM with
// their explicit inverse.  The inverses grow from the 128-wide ones by doubling,
//     inv([A B; 0 C]) = [inv(A)  -inv(A) B inv(C); 0  inv(C)],
// where every level is two grouped tcgen05 GEMMs over ALL block pairs of ALL layers at once (block-pair tile pattern of
// gemm_tc_kernel): T = -(Q Z) on the (even, odd) blocks, then Z(even, odd) = Z T.  A larger base trades ~2 n base^2 / 3
// extra flops per factor for far fewer, far larger launches: with 128-wide bases the 63 launches of a 4096-wide solve
// spent 2/3 of their time in GEMMs with K <= 256 that run at 20-40 TFLOP/s.
static int trsm_base(const psgd_ctx* ctx) { return ctx->opt_trsm_base; }

static int build_block_inverses(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int n) {
  const int base = trsm_base(ctx);
  PSGD_RETURN_IF(invert_diag_blocks(ctx, ts, count, ldq, n, n));
  std::vector<la::Gemm> gs(count);
  const size_t toff = (size_t)round_up(n, kInvBlock) * round_up(n, kInvBlock);
  for (int b = kInvBlock; b < base && b < n; b *= 2) {
    for (int t = 0; t < count; ++t) {       // T(k, k+1) = -(Q(k, k+1) Z(k+1, k+1)),  k even
      la::Gemm& g = gs[t];
      g = la::Gemm{};
      g.M = n; g.N = n; g.K = n;
      g.A = ts[t].Q; g.lda = ldq;
      g.B = ts[t].zinv; g.ldb = n;
      g.C = ts[t].zinv + toff; g.ldc = n;
      g.b_tri = 1; g.pair_b = b; g.pair_kind = 1; g.negate = true;
    }
    PSGD_RETURN_IF(gemm_many(ctx, gs.data(), count, true));
    for (int t = 0; t < count; ++t) {       // Z(k, k+1) = Z(k, k) T(k, k+1)
      la::Gemm& g = gs[t];
      g = la::Gemm{};
      g.M = n; g.N = n; g.K = n;
      g.A = ts[t].zinv; g.lda = n;
      g.B = ts[t].zinv + toff; g.ldb = n;
      g.C = ts[t].zinv; g.ldc = n;
      g.a_tri = 1; g.pair_b = b; g.pair_kind = 2;
    }
    PSGD_RETURN_IF(gemm_many(ctx, gs.data(), count, true));
  }
  return PSGD_OK;
}

static int split_point(int lo, int hi, int base) {
  const int half = (hi - lo) / 2;
  return lo + ((half + base - 1) / base) * base;
}

// Solved columns live in X (the result), columns still to be solved in W (a working copy of the right-hand side):
//   leaf    X[:, j0:j1]  = W[:, j0:j1] Z[j0:j1, j0:j1]            (out of place: a base block spans several tiles)
//   update  W[:, mid:j1] -= X[:, j0:mid] Q[j0:mid, mid:j1]
static int trsm_right_rec(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldx, int m, int n, int j0, int j1) {
  std::vector<la::Gemm> gs(count);
  if (j1 - j0 <= trsm_base(ctx)) {
    const int jb = j1 - j0;
    for (int t = 0; t < count; ++t) {
      la::Gemm& g = gs[t];
      g = la::Gemm{};
      g.M = m; g.N = jb; g.K = jb;
      g.A = ts[t].work + j0; g.lda = ldx;
      g.B = ts[t].zinv + (size_t)j0 * n + j0; g.ldb = n;
      g.C = ts[t].X + j0; g.ldc = ldx;
      g.b_tri = 1;
    }
    return gemm_many(ctx, gs.data(), count, true);
  }
  const int mid = split_point(j0, j1, trsm_base(ctx));
  PSGD_RETURN_IF(trsm_right_rec(ctx, ts, count, ldq, ldx, m, n, j0, mid));
  for (int t = 0; t < count; ++t) {
    la::Gemm& g = gs[t];
    g = la::Gemm{};
    g.M = m; g.N = j1 - mid; g.K = mid - j0;
    g.A = ts[t].X + j0; g.lda = ldx;
    g.B = ts[t].Q + (size_t)j0 * ldq + mid; g.ldb = ldq;
    g.C = ts[t].work + mid; g.ldc = ldx; g.D = g.C; g.ldd = ldx;
  }
  PSGD_RETURN_IF(gemm_many(ctx, gs.data(), count, true));
  return trsm_right_rec(ctx, ts, count, ldq, ldx, m, n, mid, j1);
}

//   leaf    X[i0:i1, :]  = Z[i0:i1, i0:i1]^T W[i0:i1, :]
//   update  W[mid:i1, :] -= Q[i0:mid, mid:i1]^T X[i0:mid, :]
static int trsm_left_rec(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldx, int m, int n, int i0, int i1) {
  std::vector<la::Gemm> gs(count);
  if (i1 - i0 <= trsm_base(ctx)) {
    const int ib = i1 - i0;
    for (int t = 0; t < count; ++t) {
      la::Gemm& g = gs[t];
      g = la::Gemm{};
      g.M = ib; g.N = m; g.K = ib;
      g.A = ts[t].zinv + (size_t)i0 * n + i0; g.lda = n; g.ta = true;
      g.B = ts[t].work + (size_t)i0 * ldx; g.ldb = ldx;
      g.C = ts[t].X + (size_t)i0 * ldx; g.ldc = ldx;
      g.a_tri = 2;
    }
    return gemm_many(ctx, gs.data(), count, true);
  }
  const int mid = split_point(i0, i1, trsm_base(ctx));
  PSGD_RETURN_IF(trsm_left_rec(ctx, ts, count, ldq, ldx, m, n, i0, mid));
  for (int t = 0; t < count; ++t) {
    la::Gemm& g = gs[t];
    g = la::Gemm{};
    g.M = i1 - mid; g.N = m; g.K = mid - i0;
    g.A = ts[t].Q + (size_t)i0 * ldq + mid; g.lda = ldq; g.ta = true;
    g.B = ts[t].X + (size_t)i0 * ldx; g.ldb = ldx;
    g.C = ts[t].work + (size_t)mid * ldx; g.ldc = ldx; g.D = g.C; g.ldd = ldx;
  }
  PSGD_RETURN_IF(gemm_many(ctx, gs.data(), count, true));
  return trsm_left_rec(ctx, ts, count, ldq, ldx, m, n, mid, i1);
}

static bool trsm_tc_ok(const psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldx, int n, int m) {
  if (ctx->opt_gemm_path == 1) return false;
  const bool big = n >= 512 && m >= 256;
  if (!(ctx->opt_gemm_path == 2 || big)) return false;
  if (n <= kInvBlock || (n % 4) != 0 || (ldq % 4) != 0 || (ldx % 4) != 0) return false;
  for (int t = 0; t < count; ++t)
    if (!aligned16(ts[t].Q) || !aligned16(ts[t].X) || !ts[t].zinv || !ts[t].work || !aligned16(ts[t].work)) return false;
  return true;
}

// X = B Q^-1 : Q [n,n] upper, B,X [m,n]
int trsm_right_many(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldb, int ldx, int m, int n) {
  if (count <= 0) return PSGD_OK;
  if (!trsm_tc_ok(ctx, ts, count, ldq, ldx, n, m)) {
    for (int t = 0; t < count; ++t)
      PSGD_RETURN_IF(la::trsm_right_upper(ctx, ts[t].Q, ldq, ts[t].B, ldb, ts[t].X, ldx, m, n));
    return PSGD_OK;
  }
  for (int t = 0; t < count; ++t)
    PSGD_CUDA_CHECK(cudaMemcpy2DAsync(ts[t].work, (size_t)ldx * 4, ts[t].B, (size_t)ldb * 4, (size_t)n * 4, m,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
  PSGD_RETURN_IF(build_block_inverses(ctx, ts, count, ldq, n));
  return trsm_right_rec(ctx, ts, count, ldq, ldx, m, n, 0, n);
}

// X = Q^-T B : Q [n,n] upper, B,X [n,m]
int trsm_left_many(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldb, int ldx, int n, int m) {
  if (count <= 0) return PSGD_OK;
  if (!trsm_tc_ok(ctx, ts, count, ldq, ldx, n, m)) {
    for (int t = 0; t < count; ++t)
      PSGD_RETURN_IF(la::trsm_left_upper_adjoint(ctx, ts[t].Q, ldq, ts[t].B, ldb, ts[t].X, ldx, n, m));
    return PSGD_OK;
  }
  for (int t = 0; t < count; ++t)
    PSGD_CUDA_CHECK(cudaMemcpy2DAsync(ts[t].work, (size_t)ldx * 4, ts[t].B, (size_t)ldb * 4, (size_t)m * 4, n,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
  PSGD_RETURN_IF(build_block_inverses(ctx, ts, count, ldq, n));
  return trsm_left_rec(ctx, ts, count, ldq, ldx, m, n, 0, n);
}

}  // namespace tc
}  // namespace psgd

// Diagnostic / building-block entry point: C = op(A) op(B) through a chosen engine (tests compare the engines).
extern "C" int psgd_gemm(psgd_ctx* ctx, int engine, int M, int N, int K, const float* A, int lda, int ta,
                         const float* B, int ldb, int tb, float* C, int ldc, int triu, int a_tri, int b_tri) {
  using namespace psgd;
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(A && B && C, PSGD_ERR_BAD_POINTER, "psgd_gemm: null device pointer");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  la::Gemm g;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.ta = ta != 0; g.B = B; g.ldb = ldb; g.tb = tb != 0;
  g.C = C; g.ldc = ldc; g.triu = triu != 0; g.a_tri = a_tri; g.b_tri = b_tri;
  if (engine == 1) return la::gemm_simt(ctx, g);
  if (engine == 2) return tc::gemm_tc(ctx, g);
  return tc::gemm_auto(ctx, g);
}
