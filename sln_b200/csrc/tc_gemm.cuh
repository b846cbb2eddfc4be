// tcgen05 / TMEM contraction core (sm_100a):  C[i,j] = sum_k A(i,k) * B(j,k)  in "3xTF32" arithmetic.
//
// Why 3xTF32: the parity contract of this path is fp32 (single-pass TF32 puts boxes_pred 5e-2 off,
// 3xTF32 stays at the fp32 noise floor).  Every fp32 operand x is split on the fly into hi = tf32(x) and lo = tf32(x - hi)
// and the tensor core accumulates  lo*hi + hi*lo + hi*hi  in fp32 in TMEM (short chains only, see "Accumulation scheme").
//
// Structure of one CTA (544 threads, tile 128 x BN, K consumed in chunks of 32 floats = one 128-byte swizzle row):
//   16 PRODUCER warps in two groups of 8 that take the chunks alternately.  A group reads A (and B, unless it is a pre-split
//     image) through the same operand functors as the SIMT kernel (gather+concat, lazy BatchNorm+ReLU, BN-backward dy,
//     transposed reads, im2col ...), splits hi/lo in registers and writes the K-major SWIZZLE_128B tiles (A_hi, A_lo, B_hi,
//     B_lo) of a shared-memory stage, then publishes it (fence.proxy.async + mbarrier arrive); pre-split weight images
//     (PackedB) arrive by cp.async.bulk onto the same mbarrier;
//   a seventeenth warp is the MMA ISSUER: per stage 4 k-slices x 2 tcgen05.mma.kind::tf32 (M=128, N=2BN and N=BN, K=8) into
//     TMEM accumulators, tcgen05.commit hands the stage back.  A ring of 3-4 stages decouples the two sides;
//   EPILOGUE (the producer warps): tcgen05.ld 32x32b (thread = accumulator row, 32 or 16 columns at a time, the columns split
//     over the warps' column groups) -> staging in shared memory -> epilogue functor with lanes along the columns (bias /
//     ReLU mask / BatchNorm column statistics / RED.ADD / the paired SPADE modulation).
// TMA tensor maps are not used for the activations because every such operand needs a per-element transform (gather, BN, hi/lo split) between
// global memory and the tensor core; the stores are laid out so that each st.shared.v4 phase covers one full 128-byte row.
#pragma once
#include <stdlib.h>

#include "gemm.cuh"

namespace sln {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;            // floats per K chunk: 128 bytes = one SWIZZLE_128B row
#ifndef SLN_TC_PF_DEEP
#define SLN_TC_PF_DEEP 2      // chunks in flight per CTA (register prefetch) for single-load A functors with pre-split B
#endif
#ifndef SLN_TC_PROD_WARPS
#define SLN_TC_PROD_WARPS 16
#endif
// producer / epilogue warps.  16 (4 per SM sub-partition): with 8 the producer instruction stream (operand transform, hi/lo split,
// index arithmetic: ~2000 warp-instructions per 32-k chunk) ran 2 warps per scheduler and was latency-bound — ncu on the largest
// SPADE kernel: issue slots 32 % busy, top stalls `wait` / `long_scoreboard`, tensor pipe 39 % (profiles/r2_prof_tc_spade_*).
constexpr int PROD_WARPS = SLN_TC_PROD_WARPS;
constexpr int PROD_THREADS = PROD_WARPS * 32;
// Producer groups: the producer warps split into PROD_GROUPS groups that take the chunks round-robin (group g: chunks g, g + G, ...),
// each group building whole stages on its own.  One chunk is a serial chain per warp (wait empty -> consume the prefetched
// registers -> st.shared -> proxy fence -> arrive -> issue the next loads): ~1400 cycles measured against 800 cycles of MMA work,
// and with every warp on the same chunk nothing overlaps it.  Two groups keep two such chains in flight.
// Measured (B200): SPADEGenerator4 419 -> 435 images/s, VAE train step 3.00 -> 2.82 ms (G = 2, one chunk prefetched per group;
// two per group: 2.86 ms).
#ifndef SLN_TC_PROD_GROUPS
#define SLN_TC_PROD_GROUPS 2
#endif
constexpr int PROD_GROUPS = SLN_TC_PROD_GROUPS;
static_assert(PROD_WARPS % PROD_GROUPS == 0 && (PROD_WARPS / PROD_GROUPS) % 4 == 0, "producer groups must be whole multiples of 4 warps");
constexpr int MMA_WARP = PROD_WARPS;        // the warp after the producers issues the tcgen05.mma stream
constexpr int THREADS = PROD_THREADS + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes (rows) x 32 consecutive fp32 columns of the accumulator: thread t gets row (lane quadrant base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in [0,14), leading byte
// offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
//   K-major tile  [rows][32 floats]: canonical ((8,n),2):((8,SBO),1) in 16-byte units — 8-row groups 1024 B apart (SBO), LBO unused;
//                 one MMA (K = 8) reads 32 bytes of every row: the k-slice advances the start address by 32 B.
//   MN-major tile [32 k][rows]:      TF32 MN-major operands must use SWIZZLE_128B_BASE32B (layout type 1): canonical
//                 ((8,n),(4,k)):((1,LBO),(8,SBO)) — a 512-byte atom holds 4 k-rows of 32 consecutive rows (128 B each) and the
//                 32-byte chunk index (address bits 5-6) is XORed with k mod 4 (bits 7-8).  Here the 8 k-atoms of a 32-row column
//                 block are contiguous (SBO = 512 B) and column blocks follow each other (LBO = 4096 B); one MMA (K = 8) reads
//                 two k-atoms: the k-slice advances the start address by 1024 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
template <bool RC>
__device__ __forceinline__ uint64_t make_desc_t(uint32_t tile, int kslice) {
  return RC ? make_desc(tile + kslice * 32) : make_desc_mn(tile + kslice * 1024);
}
// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A/B tf32 (bits 7-9, 10-12 = 2),
// a_major bit 15 / b_major bit 16 (0 = K-major, 1 = MN-major), N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

// hi = x rounded to TF32 (10 explicit mantissa bits), round-to-nearest with ties away from zero in the magnitude: an integer
// add + mask on the bit pattern (two full-rate ALU ops; identical to cvt.rna.tf32.f32 for finite x, Inf stays Inf)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// byte offset of element (row, k) inside a [rows][32-float] SWIZZLE_128B K-major tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128(int row, int k) { return (uint32_t)(row * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2))); }
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void split_store4(uint32_t hi_addr, uint32_t lo_addr, float4 v) {
  const float hx = tf32_hi(v.x), hy = tf32_hi(v.y), hz = tf32_hi(v.z), hw = tf32_hi(v.w);
  sts4(hi_addr, hx, hy, hz, hw);
  sts4(lo_addr, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
}
// Pre-split, pre-tiled B operand (weights that do not change between launches: convolution kernels at inference, Linear weights
// between optimizer steps).  sln_pack_weights() writes, for every (32-row group, 32-k chunk) UNIT, the K-major SWIZZLE_128B image of
// the hi tile (4096 B) followed by the lo tile (4096 B).  The kernel then moves a stage's B tiles with cp.async.bulk (the TMA
// engine's linear mode, SASS UBLKCP) straight into shared memory — no registers, no producer instructions, completion counted on the
// stage's `full` mbarrier — instead of loading, splitting and storing them with the producer warps.
struct PackedB {
  static constexpr bool kTwoLoads = false;
  const float* units;   // [ceil(N/128)*4][ceil(K/32)][2][32][32] floats
  int kchunks;          // ceil(K/32)
  // functor API stubs (never called: the kernel takes the bulk-copy path for this type)
  struct Tok { int unused; };
  __device__ __forceinline__ Tok token(int) const { return Tok{0}; }
  __device__ __forceinline__ int clampc(int c) const { return c; }
  __device__ __forceinline__ void fetch4(const Tok&, int, float4&, float4&) const {}
  __device__ __forceinline__ float4 finish4(const Tok&, int, float4 v, float4) const { return v; }
  bool vec_ok() const { return ((uintptr_t)units % 16) == 0; }
};
template <class T> struct is_packed { static constexpr bool value = false; };
template <> struct is_packed<PackedB> { static constexpr bool value = true; };
constexpr int PACK_UNIT_FLOATS = 2 * 32 * 32;   // hi + lo of a 32 x 32 unit
inline size_t packed_weight_floats(int64_t N, int64_t K) { return (size_t)(ceil_div64(N, 128) * 4) * (size_t)ceil_div64(K, 32) * PACK_UNIT_FLOATS; }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

// Operand loader.  Every access is a float4 of 4 consecutive STORAGE columns of one storage row, fetched through the
// functor's two-phase API: fetch() issues all raw 16-byte loads of the chunk back to back (nothing depends on them, so
// they are all in flight together), store() applies the lazy transform, splits hi/lo and writes the swizzled tiles.
// Everything that does not change from chunk to chunk (row pointers, shared-memory offsets) is computed once in init().
//   RC == true : storage [row][k] -> K-major tile.  Thread (tid >> 3, tid & 7) owns the k-quad tid & 7 of tile rows
//                (tid >> 3) + 32 i; 8 threads cover one 128-byte row chunk (coalesced), each st.shared.v4 quarter-warp
//                covers one full swizzle row (conflict-free).
//   RC == false: storage [k][row] -> MN-major tile (the tensor core transposes, not the loader).  Thread (tid >> 3, tid & 7)
//                owns, at k-row tid >> 3, the 4 consecutive tile rows 32 i + 4 (tid & 7): again 8 threads read one 128-byte
//                segment of a storage row and write one full 128-byte swizzle row.
// Bounds: tile rows beyond the operand are clamped by the functor (their products are never stored); only the K padding of
// the last chunk (CHECK == true) is zeroed.
template <int ROWS, bool RC, class Op, int NT>
struct LoaderBuf {   // the registers holding one fetched chunk (PF of these per operand)
  static constexpr int NVR = ROWS * BK / 4 / NT;  // quads per thread per chunk (0: the tile has fewer quads than threads)
  static constexpr int NV = NVR > 0 ? NVR : 1;
  float4 ra[NV], rb[NV];
  typename Op::Tok ktok;               // RC == false: the token of this chunk's k-row
};
template <int ROWS, bool RC, class Op, int NT>      // NT: threads that build one tile (a producer group)
struct Loader {      // per-thread loop invariants, shared by all register buffers
  static constexpr int NVR = ROWS * BK / 4 / NT;
  static constexpr int NV = NVR > 0 ? NVR : 1;
  static constexpr int ACTIVE = NVR > 0 ? NT : ROWS * BK / 4;   // threads that own a quad (small B tiles: the first ROWS * 8)
  static constexpr int RPP = ACTIVE / 8;               // RC: tile rows covered per pass (quad i: + RPP rows)
  static constexpr int KB = ACTIVE / 256 > 0 ? ACTIVE / 256 : 1;   // !RC: 32-row blocks covered per pass by the 32 k-rows x 8 row-quads
  using Buf = LoaderBuf<ROWS, RC, Op, NT>;
  typename Op::Tok tok[RC ? NV : 1];   // RC: one token per owned tile row
  int col[RC ? 1 : NV];                // !RC: clamped storage column of quad i
  uint32_t soff;                       // byte offset of quad 0 inside the tile
  bool active;
  __device__ __forceinline__ void init(const Op& op, int row0, int tid) {
    active = tid < ACTIVE;
    const int t = active ? tid : 0;
    if (RC) {
      soff = sw128(t >> 3, (t & 7) * 4);                                   // K-major SWIZZLE_128B: (row, k-quad); quad i: + RPP rows
#pragma unroll
      for (int i = 0; i < NV; ++i) tok[i] = op.token(row0 + (t >> 3) + RPP * i);
    } else {
      // MN-major 128B_BASE32B: thread = (k-row kr = (t >> 3) & 31, row-quad t & 7, 32-row block t >> 8); quad i: block + KB i
      const int kr = (t >> 3) & 31, blk = t >> 8;
      soff = (uint32_t)(kr * 128 + (((((t & 7) >> 1) ^ (kr & 3)) << 5) | ((t & 1) << 4))) + (uint32_t)blk * 4096u;
#pragma unroll
      for (int i = 0; i < NV; ++i) col[i] = op.clampc(row0 + 32 * (blk + KB * i) + 4 * (t & 7));
    }
  }
  static constexpr uint32_t QSTEP = RC ? (uint32_t)RPP * 128u : (uint32_t)KB * 4096u;   // tile bytes between a thread's consecutive quads
  template <bool CHECK>
  __device__ __forceinline__ void fetch(const Op& op, Buf& b, int k0, int kend, int tid) const {
    if (!active) return;
    if (RC) {
      int k = k0 + (tid & 7) * 4;
      if (CHECK) k = min(k, kend - 4);
#pragma unroll
      for (int i = 0; i < NV; ++i) op.fetch4(tok[i], k, b.ra[i], b.rb[i]);
    } else {
      int k = k0 + ((tid >> 3) & 31);
      if (CHECK) k = min(k, kend - 1);
      b.ktok = op.token(k);
#pragma unroll
      for (int i = 0; i < NV; ++i) op.fetch4(b.ktok, col[i], b.ra[i], b.rb[i]);
    }
  }
  // (k0, kend, CHECK) must be the ones passed to the matching fetch()
  template <bool CHECK>
  __device__ __forceinline__ void store(const Op& op, const Buf& b, int k0, int kend, uint32_t hi_tile, uint32_t lo_tile, int tid) const {
    if (!active) return;
    if (RC) {
      const int k = k0 + (tid & 7) * 4;
      const bool valid = !CHECK || k < kend;
      const int kc = CHECK ? min(k, kend - 4) : k;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 v = op.finish4(tok[i], kc, b.ra[i], b.rb[i]);
        if (CHECK && !valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
        split_store4(hi_tile + soff + i * QSTEP, lo_tile + soff + i * QSTEP, v);
      }
    } else {
      const bool valid = !CHECK || (k0 + ((tid >> 3) & 31)) < kend;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 v = op.finish4(b.ktok, col[i], b.ra[i], b.rb[i]);
        if (CHECK && !valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
        split_store4(hi_tile + soff + i * QSTEP, lo_tile + soff + i * QSTEP, v);
      }
    }
  }
};

// ---------------------------------------------------------------- epilogues
// The accumulator tile is staged through shared memory (row-per-thread float4 writes, conflict-free with a BN+4 row
// stride) and then processed with lanes along the columns, so every global access of the epilogue is a coalesced float4
// and BatchNorm column statistics are plain per-thread sums.  apply4() handles 4 consecutive columns j..j+3 (< N, row < M
// guaranteed by the caller for full quads; `nvalid` = number of valid columns in the quad).
struct TcEpiStore {   // C = acc + bias (+ BatchNorm batch statistics)       reference graph.py:12-15
  float* C; int ldc; const float* bias; BnFwdFin fin;
  static constexpr bool kStats = true;
  static constexpr bool kPaired = false;
  __device__ __forceinline__ bool wants_stats() const { return fin.enabled != 0; }
  __device__ __forceinline__ void apply4(int i, int j, int nvalid, float4 a, float (&s1)[4], float (&s2)[4]) const {
    float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < nvalid) {
        if (bias) v[e] += __ldg(bias + j + e);
        s1[e] += v[e];
        s2[e] = fmaf(v[e], v[e], s2[e]);
      }
    }
    float* dst = C + (size_t)i * ldc + j;
    if (nvalid == 4 && ((uintptr_t)dst % 16 == 0)) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (e < nvalid) dst[e] = v[e];
    }
  }
  __device__ __forceinline__ void finalize(int col, double, double, double S, double Q, int Mt) const { bn_fwd_apply(fin, col, S, Q, Mt); }
  __device__ __forceinline__ float* partial() const { return fin.partial; }
  __device__ __forceinline__ unsigned* counter() const { return fin.counter; }
  __device__ __forceinline__ const sln_bn_sync* sync() const { return fin.sync; }
  __device__ __forceinline__ int slot0() const { return fin.slot0; }
  __device__ __forceinline__ int rows() const { return fin.M; }
};

struct TcEpiMaskReduce {   // G = relu_mask(yprev) ? (acc + add) : 0  + BN-backward column sums
  float* G; int ldg; const float* add; int ldadd; const float* yprev; int ldy;
  const float* scale; const float* shift; const float* mean; const float* rstd; BnBwdFin fin;
  static constexpr bool kStats = true;
  static constexpr bool kPaired = false;
  __device__ __forceinline__ bool wants_stats() const { return true; }
  __device__ __forceinline__ void apply4(int i, int j, int nvalid, float4 a, float (&s1)[4], float (&s2)[4]) const {
    float d[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < nvalid) {
        const int jj = j + e;
        float y = __ldg(yprev + (size_t)i * ldy + jj);
        float pre = scale ? fmaf(y, __ldg(scale + jj), __ldg(shift + jj)) : y;
        float dd = d[e];
        if (add) dd += __ldg(add + (size_t)i * ldadd + jj);
        float g = pre > 0.f ? dd : 0.f;
        G[(size_t)i * ldg + jj] = g;
        s1[e] += g;
        float yh = mean ? (y - __ldg(mean + jj)) * __ldg(rstd + jj) : 0.f;
        s2[e] = fmaf(g, yh, s2[e]);
      }
    }
  }
  __device__ __forceinline__ void finalize(int col, double S, double Q, double Sa, double Qa, int Mt) const { bn_bwd_apply(fin, col, S, Q, Sa, Qa, Mt); }
  __device__ __forceinline__ float* partial() const { return fin.partial; }
  __device__ __forceinline__ unsigned* counter() const { return fin.counter; }
  __device__ __forceinline__ const sln_bn_sync* sync() const { return fin.sync; }
  __device__ __forceinline__ int slot0() const { return fin.slot0; }
  __device__ __forceinline__ int rows() const { return fin.M; }
};

struct TcEpiAtomic {   // C += acc (split-K weight gradients, RED.ADD)
  float* C; int ldc;
  static constexpr bool kStats = false;
  static constexpr bool kPaired = false;
  __device__ __forceinline__ bool wants_stats() const { return false; }
  __device__ __forceinline__ void apply4(int i, int j, int nvalid, float4 a, float (&s1)[4], float (&s2)[4]) const {
    float* dst = C + (size_t)i * ldc + j;
    if (nvalid == 4 && ((uintptr_t)dst % 16 == 0)) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    } else {
      float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) if (e < nvalid) red_add(dst + e, v[e]);
    }
  }
  __device__ __forceinline__ void finalize(int, double, double, double, double, int) const {}
  __device__ __forceinline__ float* partial() const { return nullptr; }
  __device__ __forceinline__ unsigned* counter() const { return nullptr; }
  __device__ __forceinline__ const sln_bn_sync* sync() const { return nullptr; }
  __device__ __forceinline__ int slot0() const { return 0; }
  __device__ __forceinline__ int rows() const { return 0; }
};

template <int BN>
struct SmemLayout {
  static constexpr int A_TILE = BM * 128;   // bytes of one [128][32] fp32 tile
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = BN >= 128 ? 3 : 4;     // 192 KB / 192 KB / 160 KB of operand ring
  static constexpr int STAT = 2 * PROD_WARPS * BN * 4;  // per-warp column statistics [2][PROD_WARPS][BN]
  static constexpr int OUT_LD = BN + 4;                // floats per staged accumulator row
  static_assert(BM * OUT_LD * 4 <= STAGES * STAGE, "accumulator staging must fit the (idle) stage buffers");
  static constexpr int NBARS = 2 * STAGES + 4;         // full[S] empty[S] segfull[2] accempty[2]
  static constexpr int BYTES = 1024 + STAGES * STAGE + STAT + NBARS * 8 + 64;
};

// Accumulation scheme.
//   * Dependent tcgen05.mma instructions (same TMEM accumulator) serialise on the accumulate latency (~170 cycles measured),
//     far above the 16-64 cycle occupancy of a 128 x BN x 8 TF32 instruction.  The products of the 3xTF32 scheme are
//     therefore kept in SEPARATE accumulators ([hi*hi | hi*lo] from one N = 2 BN instruction, lo*hi from a second one) and,
//     for BN <= 64, even and odd k-slices alternate between two accumulator sets: 2-4 independent MMA chains per CTA.
//   * The tensor core adds into its fp32 accumulator with truncation, so a long dependent chain of accumulator updates
//     drifts (measured: 1.4e-5 relative after 1488 updates).  Chains are bounded by draining all accumulators into fp32
//     registers (round-to-nearest adds) every SEG_CHUNKS chunks = 1280 k (80 updates per chain, < 1e-6).  Contractions whose
//     K range per CTA fits one segment (every contraction of the VAE path, most convolutions) run the MSEG == false variant:
//     no in-loop drain, so the register accumulators only exist in the epilogue and two chunks are prefetched per operand.
// TMEM columns: per accumulator set [hi*hi | hi*lo | lo*hi], BN columns each (one set for BN = 128, an even and an odd set below).
constexpr int SEG_CHUNKS = 40;

// tuning aid: SM-clock timestamps of CTA (0,0,0) at the phase boundaries of the last tc_gemm launch (sln_debug_tc_trace)
__device__ long long g_tc_trace[48];
#ifdef SLN_TC_TRACE
#define TC_TRACE(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 31) == 0 && (tid == 0 || tid == MMA_WARP * 32)) g_tc_trace[(slot) + (tid ? 8 : 0)] = clock64(); } while (0)
#define TC_TRACE2(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || tid == MMA_WARP * 32)) g_tc_trace[16 + (slot) + (tid ? 16 : 0)] = clock64(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#define TC_TRACE2(slot) do { } while (0)
#endif

// One lane of a CONVERGED warp, chosen by the hardware (elect.sync).  ptxas knows the elected lane is unique, so tcgen05.mma /
// cp.async.bulk / tcgen05.commit issued under it compile to straight-line UTCHMMA / UBLKCP / UTCBAR with uniform-register operands.
// Under a plain `if (lane == 0)` the compiler cannot prove uniformity and wraps EVERY such instruction in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop (~100 cycles per MMA, measured: the issue loop, not the tensor core, paced
// the BN <= 64 tiles).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One CTA = PROD_WARPS producer/epilogue warps (in PROD_GROUPS groups) + 1 MMA warp, decoupled by mbarriers (no CTA-wide barrier in the main loop):
//   producers   wait empty[s] -> transform + hi/lo split + st.shared of chunk c (fetched PF chunks earlier into registers)
//               -> issue the global loads of chunk c+PF -> fence.proxy.async -> one arrive per warp on full[s];
//   MMA warp    (one lane) wait full[s] -> 12 tcgen05.mma.kind::tf32 -> tcgen05.commit -> empty[s]  (-> segfull at segment ends);
//   the ring of STAGES stages lets the producers run ahead of the tensor core, so global-load latency is hidden behind
//   the MMAs of the previous chunks instead of being exposed once per chunk.
// grid = (ceil(N/BN), ceil(M/128), splits); each z-slice reduces k in [z*kchunk, (z+1)*kchunk), kchunk % 32 == 0.
template <int BN, bool A_RC, bool B_RC, bool MSEG, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const AOp A, const BOp B, const Epi epi, int M, int N, int K, int kchunk, int xmap) {
  using L = SmemLayout<BN>;
  // chunks prefetched into registers per producer thread.  The loads of chunk c + PF are issued at the end of produce(c) and consumed
  // at the start of produce(c + PF), i.e. PF - 1 chunk periods later: with PF = 2 the period settled at the (loaded) L2 latency of
  // ~900 cycles (phase trace, profiles/r2_trace_contract.txt).  Operands that arrive as a pre-split image (PackedB) need no B
  // registers, so the single-load A functors can afford 4 chunks in flight (64 registers; one CTA per SM allows 224).
  constexpr int S = L::STAGES, PF = MSEG ? 1 : ((is_packed<BOp>::value && !AOp::kTwoLoads) ? SLN_TC_PF_DEEP : 2);
  constexpr bool BP = is_packed<BOp>::value;       // B tiles arrive by cp.async.bulk from a pre-split, pre-tiled image
  // producer groups; the 128-column multi-segment variant has no registers for a second chunk in flight (96-register cap: 600 B of spills)
  constexpr int G = (MSEG && BN >= 128) ? 1 : PROD_GROUPS, GROUP_WARPS = PROD_WARPS / G, GROUP_THREADS = GROUP_WARPS * 32;
  constexpr int PFG = PF / G > 0 ? PF / G : 1;     // chunks prefetched per thread of a group (PFG * G chunks in flight per CTA)
  extern __shared__ char smem_raw[];
  char* smem = (char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles need 1024-byte alignment
  float* stat = reinterpret_cast<float*>(smem + S * L::STAGE);             // [2][PROD_WARPS][BN]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * L::STAGE + L::STAT);
  uint64_t* empty = full + S;
  uint64_t* segfull = empty + S;
  uint64_t* accempty = segfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // xmap != 0: the launch covers a subset of the column tiles — tile = blockIdx.x, plus (xmap >> 16) from tile (xmap & 0xffff) on
  const int xtile = (int)blockIdx.x + (((xmap >> 16) != 0 && (int)blockIdx.x >= (xmap & 0xffff)) ? (xmap >> 16) : 0);
  const int m0 = blockIdx.y * BM, n0 = xtile * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int nchunks = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;
  TC_TRACE(0);
  // Programmatic dependent launch: let the next kernel of the stream start its prologue (TMEM allocation, barrier set-up) on idle
  // SMs while this grid runs; that kernel's own griddepcontrol.wait (below) orders its first global read after our last write.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // Accumulator regions (BN columns each).  The hi and lo tiles of B are adjacent in shared memory, so ONE instruction with N = 2 BN
  // computes A_hi x [B_hi ; B_lo]^T = [hi*hi | hi*lo]; a second one (N = BN) adds A_lo x B_hi^T = lo*hi: 2 tcgen05.mma per k-slice
  // instead of 3, and A_hi is read from shared memory once.  BN <= 64 keeps two such sets (even / odd k-slices: independent chains).
  constexpr int NSET = BN >= 128 ? 1 : 2;
  constexpr int NREG = 3 * NSET;
  constexpr uint32_t TMEM_COLS = (NREG * BN <= 128) ? 128 : ((NREG * BN <= 256) ? 256 : 512);
  static_assert(NREG * BN <= 512, "the accumulators must fit the 512 TMEM columns");

  if (warp == MMA_WARP) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) { mbar_init(full + s, GROUP_WARPS + (BP ? 1 : 0)); mbar_init(empty + s, 1); }
      mbar_init(segfull, 1); mbar_init(segfull + 1, 1);
      mbar_init(accempty, PROD_WARPS); mbar_init(accempty + 1, PROD_WARPS);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // everything above overlapped the previous kernel's tail; its results are visible from here on
  TC_TRACE(1);

  const int quad = warp & 3, cg = warp >> 2;            // TMEM lane quadrant (fixed by warp id % 4), column group
  // the BN accumulator columns are split over the PROD_WARPS/4 column groups in chunks of CW = 32 (or 16) columns: with 8 warps
  // and BN = 128 a thread owns 2 x 32 columns, with 16 warps 32 (BN = 128) or 16 (BN = 64): half the registers of the
  // multi-segment running sum and half the TMEM reads per warp; groups past BN/CW idle
  constexpr int NCG = PROD_WARPS / 4, CW = BN / NCG >= 32 ? 32 : 16;   // column groups, chunk width
  constexpr int CGA = BN / CW < NCG ? BN / CW : NCG;    // groups that hold accumulators
  constexpr int CH2 = BN / CW / CGA;                    // CW-column chunks owned by one thread
  const bool has_acc = cg < CGA;                        // all producer warps apply the epilogue, these read the accumulators
  const int col_off = cg * (BN / CGA);
  const uint32_t tmem_mine = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)col_off;
  float racc[MSEG ? CH2 : 1][CW];      // MSEG: running sum over the drained segments (otherwise the sum is formed in the epilogue)
  if (MSEG) {
#pragma unroll
    for (int j = 0; j < (MSEG ? CH2 : 1); ++j)
#pragma unroll
      for (int e = 0; e < CW; ++e) racc[j][e] = 0.f;
  }

  if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BN, !A_RC, !B_RC), idesc2 = make_idesc(2 * BN, !A_RC, !B_RC);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S, use = c / S;
        const int seg = c / SEG_CHUNKS;
        const bool seg_first = (c % SEG_CHUNKS) == 0, seg_last = (c % SEG_CHUNKS) == SEG_CHUNKS - 1 || c == nchunks - 1;
        if (seg_first && seg > 0) {                       // the producers have drained the accumulators of the previous segment
          mbar_wait(accempty, (uint32_t)((seg - 1) & 1));
          tc_fence_after();
        }
        mbar_wait(full + s, (uint32_t)(use & 1));
        tc_fence_after();
        if (c < 8) TC_TRACE2(c);
        const uint32_t a_hi = smem_u32(smem + s * L::STAGE), a_lo = a_hi + L::A_TILE, b_hi = a_hi + 2 * L::A_TILE;   // b_lo = b_hi + B_TILE: rows BN..2BN-1 of the same tile
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {               // UMMA_K = 8 tf32: one 32-byte slice of every K-major row / one MN-major k-atom
          const uint32_t d0 = tmem_acc + (uint32_t)((k % NSET) * 3 * BN);
          const uint32_t fresh = (!seg_first || k >= NSET) ? 1u : 0u;
          mma_tf32(d0, make_desc_t<A_RC>(a_hi, k), make_desc_t<B_RC>(b_hi, k), idesc2, fresh);            // [hi*hi | hi*lo]
          mma_tf32(d0 + 2 * BN, make_desc_t<A_RC>(a_lo, k), make_desc_t<B_RC>(b_hi, k), idesc, fresh);    // lo*hi
        }
        mma_commit(empty + s);                             // stage s may be overwritten once these MMAs retire
        if (seg_last) mma_commit(segfull);
        if (c == 0) TC_TRACE(2);
      }
      TC_TRACE(3);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ producers (and owners of the register accumulators)
    auto drain = [&](int seg) {                           // MSEG only: racc += every TMEM accumulator of a retired segment
      mbar_wait(segfull, (uint32_t)(seg & 1));
      tc_fence_after();
      if (MSEG && has_acc) {
#pragma unroll
        for (int rg = 0; rg < NREG; ++rg) {
#pragma unroll
          for (int j = 0; j < (MSEG ? CH2 : 1); ++j) {
            float v[CW];
            tmem_ldw(tmem_mine + (uint32_t)(rg * BN + j * CW), v);
#pragma unroll
            for (int e = 0; e < CW; ++e) racc[j][e] += v[e];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);
    };
    // loop invariants once per operand; PFG register buffers (indexed by compile-time constants only, so that they stay in registers)
    const int grp = warp / GROUP_WARPS, gtid = tid - grp * GROUP_THREADS;   // this warp's producer group, thread index inside it
    Loader<BM, A_RC, AOp, GROUP_THREADS> la;
    Loader<BN, B_RC, BOp, GROUP_THREADS> lb;
    typename Loader<BM, A_RC, AOp, GROUP_THREADS>::Buf abuf[PFG];
    typename Loader<BN, B_RC, BOp, GROUP_THREADS>::Buf bbuf[PFG];
    const bool tail = ((kend - kbeg) & (BK - 1)) != 0;     // only the last chunk can have K padding
    const uint32_t sbase = smem_u32(smem);
    int ndrained = 0;                                      // MSEG: segments this warp has drained (every warp drains every segment, in order)
    auto fetch = [&](auto& BA, auto& BB, int c) {
      if (tail && c == nchunks - 1) {
        la.template fetch<true>(A, BA, kbeg + c * BK, kend, gtid);
        if constexpr (!BP) lb.template fetch<true>(B, BB, kbeg + c * BK, kend, gtid);
      } else {
        la.template fetch<false>(A, BA, kbeg + c * BK, kend, gtid);
        if constexpr (!BP) lb.template fetch<false>(B, BB, kbeg + c * BK, kend, gtid);
      }
    };
    auto produce = [&](auto& BA, auto& BB, int c) {
      const int s = c % S, use = c / S;
      const uint32_t st = sbase + s * L::STAGE;
      if (use > 0) mbar_wait(empty + s, (uint32_t)((use - 1) & 1));   // the MMAs that read this stage have retired
      if (c < 4) TC_TRACE2(2 * c);
      if constexpr (BP) {
        if (gtid < 32 && elect_one()) {                   // B_hi | B_lo tiles of this stage: BN/32 units, two 4 KB bulk copies each
          mbar_expect_tx(full + s, 2 * L::B_TILE);
          const float* u0 = B.units + ((size_t)(n0 / 32) * B.kchunks + (size_t)(kbeg / BK + c)) * PACK_UNIT_FLOATS;
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            const float* u = u0 + (size_t)j * B.kchunks * PACK_UNIT_FLOATS;
            bulk_g2s(st + 2 * L::A_TILE + j * 4096, u, 4096, full + s);
            bulk_g2s(st + 2 * L::A_TILE + L::B_TILE + j * 4096, u + 1024, 4096, full + s);
          }
        }
      }
      if (tail && c == nchunks - 1) {
        la.template store<true>(A, BA, kbeg + c * BK, kend, st, st + L::A_TILE, gtid);
        if constexpr (!BP) lb.template store<true>(B, BB, kbeg + c * BK, kend, st + 2 * L::A_TILE, st + 2 * L::A_TILE + L::B_TILE, gtid);
      } else {
        la.template store<false>(A, BA, kbeg + c * BK, kend, st, st + L::A_TILE, gtid);
        if constexpr (!BP) lb.template store<false>(B, BB, kbeg + c * BK, kend, st + 2 * L::A_TILE, st + 2 * L::A_TILE + L::B_TILE, gtid);
      }
      if (c < 4) TC_TRACE2(2 * c + 1);
      if (c + PFG * G < nchunks) fetch(BA, BB, c + PFG * G);   // refill the register buffer: these loads fly during the group's next PFG chunks
      fence_async_smem();                                 // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
      if (MSEG) {
        // a group's first chunk of a new segment: drain the segment before it (a group's consecutive chunks are G <= SEG_CHUNKS
        // apart, so at most one segment is pending here)
        if (c / SEG_CHUNKS > ndrained) drain(ndrained++);
      }
    };
    la.init(A, m0, gtid);
    if constexpr (!BP) lb.init(B, n0, gtid);
    TC_TRACE2(8);
#pragma unroll
    for (int i = 0; i < PFG; ++i)
      if (grp + i * G < nchunks) fetch(abuf[i], bbuf[i], grp + i * G);
    TC_TRACE2(9);
    for (int c = grp; c < nchunks; c += PFG * G) {
#pragma unroll
      for (int i = 0; i < PFG; ++i)
        if (c + i * G < nchunks) produce(abuf[i], bbuf[i], c + i * G);
    }
    if (MSEG && nchunks > 0) {
      const int last = (nchunks - 1) / SEG_CHUNKS;          // a group without a chunk in the last segment still owes the one before it
      if (G > 1 && ndrained < last) drain(ndrained++);
      drain(last);
    }
    if (!MSEG && nchunks > 0) {                           // single segment: every MMA has retired when segfull completes
      mbar_wait(segfull, 0u);
      tc_fence_after();
    }
  }
  TC_TRACE(4);
  // ---- epilogue.  Lanes run along the columns: LPR lanes cover one row (4 columns each), a warp covers 32/LPR rows per pass and
  // NIT passes in all.  Paired epilogues first issue the loads of the tensor they modulate (HBM latency: they fly during the
  // accumulator read-out and the staging barrier instead of once per pass).
  constexpr int LPR = BN / 4, RPP = 32 / LPR, NIT = BM / (PROD_WARPS * RPP);           // BN = 128: 32 lanes per row, 1 row per pass
  constexpr int HB = BN / 2, LPR2 = HB / 4, RPP2 = 32 / LPR2, NIT2 = BM / (PROD_WARPS * RPP2);
  static_assert(BM % (PROD_WARPS * RPP) == 0 && BM % (PROD_WARPS * RPP2) == 0, "whole passes only");
  const int cq2 = lane % LPR2, rsub2 = lane / LPR2;
  const int ch = n0 / 2 + cq2 * 4;                     // paired tiles: first of the 4 output channels of this lane
  float4 xin[Epi::kPaired ? NIT2 : 1];
  if constexpr (Epi::kPaired) {
    if (warp < PROD_WARPS && 2 * ch < N) {
#pragma unroll
      for (int it = 0; it < NIT2; ++it) {
        const int r = warp * RPP2 + rsub2 + it * PROD_WARPS * RPP2;
        xin[it] = m0 + r < M ? epi.load_pair(m0 + r, ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  // ... stage the tile in shared memory (the stage buffers are idle: every MMA has retired) ...
  float* outs = reinterpret_cast<float*>(smem);
  if (has_acc) {
    const int r = quad * 32 + lane;
#pragma unroll
    for (int j = 0; j < CH2; ++j) {
      float acc[CW];
      if (MSEG) {
#pragma unroll
        for (int e = 0; e < CW; ++e) acc[e] = racc[MSEG ? j : 0][e];
      } else {
#pragma unroll
        for (int e = 0; e < CW; ++e) acc[e] = 0.f;
        if (nchunks > 0) {
#pragma unroll
          for (int rg = 0; rg < NREG; ++rg) {
            float v[CW];
            tmem_ldw(tmem_mine + (uint32_t)(rg * BN + j * CW), v);
#pragma unroll
            for (int e = 0; e < CW; ++e) acc[e] += v[e];
          }
        }
      }
#pragma unroll
      for (int e = 0; e < CW; e += 4)
        *reinterpret_cast<float4*>(outs + (size_t)r * L::OUT_LD + col_off + j * CW + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
    }
  }
  TC_TRACE2(10);
  __syncthreads();
  TC_TRACE2(11);
  // ... then every warp reads its rows back (all passes first: independent shared-memory loads) and applies the epilogue
  const bool stats = Epi::kStats && epi.wants_stats();
  const int cq = lane % LPR, rsub = lane / LPR;
  const int jcol = n0 + cq * 4;
  const int nvalid = min(4, max(0, N - jcol));
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (Epi::kPaired) {
    // paired tiles: columns [0, BN/2) and [BN/2, BN) of a tile belong to the SAME BN/2 output channels (gamma | beta of a SPADE
    // modulation); a lane takes 4 channels of one row from both halves
    if (warp < PROD_WARPS && 2 * ch < N) {
      float4 ga[NIT2], be[NIT2];
#pragma unroll
      for (int it = 0; it < NIT2; ++it) {
        const int r = warp * RPP2 + rsub2 + it * PROD_WARPS * RPP2;
        ga[it] = *reinterpret_cast<const float4*>(outs + (size_t)r * L::OUT_LD + cq2 * 4);
        be[it] = *reinterpret_cast<const float4*>(outs + (size_t)r * L::OUT_LD + HB + cq2 * 4);
      }
#pragma unroll
      for (int it = 0; it < NIT2; ++it) {
        const int r = warp * RPP2 + rsub2 + it * PROD_WARPS * RPP2;
        if (m0 + r < M) epi.apply_pair(m0 + r, ch, ga[it], be[it], xin[Epi::kPaired ? it : 0]);
      }
    }
  } else if (nvalid > 0 && warp < PROD_WARPS) {
    float4 av[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) av[it] = *reinterpret_cast<const float4*>(outs + (size_t)(warp * RPP + rsub + it * PROD_WARPS * RPP) * L::OUT_LD + cq * 4);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {                   // rows in ascending order: the column statistics keep their summation order
      const int r = warp * RPP + rsub + it * PROD_WARPS * RPP;
      if (m0 + r < M) epi.apply4(m0 + r, jcol, nvalid, av[it], s1, s2);
    }
  }
  TC_TRACE(5);
  if (stats) {
    if (warp < PROD_WARPS) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {                       // fold the RPP row sub-groups of the warp (fixed order)
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) { s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o); s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o); }
      }
      if (rsub == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { stat[(0 * PROD_WARPS + warp) * BN + cq * 4 + e] = s1[e]; stat[(1 * PROD_WARPS + warp) * BN + cq * 4 + e] = s2[e]; }
      }
    }
    __syncthreads();
    float* partial = epi.partial();
    for (int j = tid; j < BN; j += THREADS) {
      if (n0 + j < N) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < PROD_WARPS; ++w) { a += stat[(0 * PROD_WARPS + w) * BN + j]; b += stat[(1 * PROD_WARPS + w) * BN + j]; }
        partial[((size_t)blockIdx.y * 2 + 0) * N + n0 + j] = a;
        partial[((size_t)blockIdx.y * 2 + 1) * N + n0 + j] = b;
      }
    }
    // (finalize_column_block starts with a CTA barrier, after which the staging area is reused as fp64 scratch)
    finalize_column_block<THREADS>(partial, epi.counter(), n0, BN, N, tid, reinterpret_cast<double*>(smem), s_last,
                                   [&](int col, double S_, double Q_, double Sa, double Qa, int Mt) { epi.finalize(col, S_, Q_, Sa, Qa, Mt); },
                                   epi.sync(), epi.slot0(), epi.rows());
  }
  TC_TRACE(6);
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_acc, TMEM_COLS);
  TC_TRACE(7);
}

// ---------------------------------------------------------------- host side
inline bool pdl_enabled() {   // env SLN_PDL=0 disables programmatic dependent launch (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("SLN_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

struct TcChoice { int bn, splits, kchunk; };
// A launch over `count` column tiles of width `bn`: the tiles 0 .. count-1, with `skip_n` tiles left out from tile `skip_at` on
// (e.g. count 4, skip_at 2, skip_n 1 -> tiles 0, 1, 3, 4; count 1, skip_at 0, skip_n 2 -> tile 2).  count == 0: every tile, width chosen
// by pick_tc.  The epilogue state (BatchNorm partials, tickets, statistics) is per column block, so disjoint subsets are independent.
struct TileSel { int count = 0, skip_at = 0, skip_n = 0, bn = 0; };

inline TcChoice pick_tc(int M, int N, int K, bool allow_split) {
  const int bns[3] = {128, 64, 32};
  double best = 1e300;
  TcChoice c{32, 1, K};
  static int forced = -1;                            // env SLN_TC_BN = 32 | 64 | 128 forces the tile width (tuning experiments)
  if (forced < 0) { const char* e = getenv("SLN_TC_BN"); forced = e ? atoi(e) : 0; }
  for (int t = 0; t < 3; ++t) {
    const int bn = bns[t];
    if (forced > 0 && bn != forced) continue;
    if (forced == 0 && bn > 32 && N <= bn / 2) continue;           // do not pad N by more than 2x
    int tiles = ceil_div(M, BM) * ceil_div(N, bn);
    int splits = 1;
    if (allow_split) {
      splits = tiles >= kNumSMs ? 1 : ceil_div(kNumSMs, tiles);
      int max_splits = ceil_div(K, 256);
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
    int kchunk = ceil_div(ceil_div(K, splits), BK) * BK;
    splits = ceil_div(K, kchunk);
    // with producer groups the 128-column multi-segment variant runs a single group (no registers for a second chunk in flight):
    // the 64-column tile with two groups is the faster multi-segment kernel (measured, K = 2304 / 9216: 2-5 %)
    if (forced == 0 && PROD_GROUPS > 1 && bn == 128 && ceil_div(kchunk, BK) > SEG_CHUNKS) continue;
    int ctas = tiles * splits;
    int waves = ceil_div(ctas, kNumSMs);            // one CTA per SM (shared-memory bound)
    // cycles per CTA ~ prologue (TMEM alloc, first loads) + chunks x max(MMA time, producer time) + epilogue:
    //   12 MMAs of 128 x bn x 8 per chunk at 128*bn/256 cycles each; producers move (128 + bn) x 32 floats per chunk
    double t_mma = 12.0 * (bn / 2 < 32 ? 32 : bn / 2), t_prod = 1.6 * (BM + bn);
    double per_cta = 2500.0 + (double)ceil_div(kchunk, BK) * (t_mma > t_prod ? t_mma : t_prod) + 1500.0 + 14.0 * bn;
    double cost = waves * per_cta;
    if (cost < best) { best = cost; c = TcChoice{bn, splits, kchunk}; }
  }
  return c;
}

// shapes the tensor-core path accepts; everything else (tiny heads, K = 6 box embedding) stays on the SIMT kernels
// (independent of M, so that a scene evaluated alone and inside a batch takes the same arithmetic path)
inline bool tc_eligible(int M, int N, int K) { return M >= 1 && N >= 32 && K >= 32; }

template <int BN, bool A_RC, bool B_RC, bool MSEG, class AOp, class BOp, class Epi>
int launch_tc_bn_seg(cudaStream_t st, const AOp& A, const BOp& B, const Epi& epi, int M, int N, int K, const TcChoice& c, const TileSel& sel) {
  auto kern = tc_gemm_kernel<BN, A_RC, B_RC, MSEG, AOp, BOp, Epi>;
  constexpr int bytes = SmemLayout<BN>::BYTES;
  static unsigned long long configured = 0ull;   // devices configured (per call site / instantiation)
  if (first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(tc_gemm, %d B smem) failed: %s", bytes, cudaGetErrorString(e)); return SLN_ECUDA; }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(sel.count > 0 ? sel.count : ceil_div(N, BN), ceil_div(M, BM), c.splits);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int kchunk = c.kchunk;
  int xmap = sel.count > 0 ? ((sel.skip_n << 16) | sel.skip_at) : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, A, B, epi, M, N, K, kchunk, xmap);
  if (e != cudaSuccess) { set_error("cudaLaunchKernelEx(tc_gemm) failed: %s", cudaGetErrorString(e)); return SLN_ECUDA; }
  return SLN_OK;
}

template <int BN, bool A_RC, bool B_RC, class AOp, class BOp, class Epi>
int launch_tc_bn(cudaStream_t st, const AOp& A, const BOp& B, const Epi& epi, int M, int N, int K, const TcChoice& c, const TileSel& sel = TileSel()) {
  // K ranges that fit one accumulation segment take the variant without in-loop drains (deeper register prefetch)
  if (ceil_div(c.kchunk < K ? c.kchunk : K, BK) > SEG_CHUNKS) return launch_tc_bn_seg<BN, A_RC, B_RC, true>(st, A, B, epi, M, N, K, c, sel);
  return launch_tc_bn_seg<BN, A_RC, B_RC, false>(st, A, B, epi, M, N, K, c, sel);
}

template <bool A_RC, bool B_RC, class AOp, class BOp, class Epi>
int launch_tc(cudaStream_t st, const AOp& A, const BOp& B, const Epi& epi, int M, int N, int K, bool allow_split, const char* what,
              int prof_cls, const TileSel& sel = TileSel()) {
  if (M <= 0 || N <= 0) return SLN_OK;
  const int ncols = sel.count > 0 ? (sel.count * sel.bn < N ? sel.count * sel.bn : N) : N;
  ProfScope prof(st, prof_cls, 2.0 * (double)M * (double)ncols * (double)K);
  TcChoice c = sel.count > 0 ? TcChoice{sel.bn, 1, ceil_div(K, BK) * BK} : pick_tc(M, N, K, allow_split);
  int rc;
  if (c.bn == 128) rc = launch_tc_bn<128, A_RC, B_RC>(st, A, B, epi, M, N, K, c, sel);
  else if (c.bn == 64) rc = launch_tc_bn<64, A_RC, B_RC>(st, A, B, epi, M, N, K, c, sel);
  else rc = launch_tc_bn<32, A_RC, B_RC>(st, A, B, epi, M, N, K, c, sel);
  if (rc != SLN_OK) return rc;
  return check_launch(what);
}

}  // namespace tc
}  // namespace sln
