// Block-diagonal irreps linear on the 5th-generation tensor cores (tcgen05 + TMEM + TMA, sm_100a):
//   C[(n,i), :N] (+)= A[(n,i), :K] * W[K, N]        fp32 in / fp32 out
// Same operand addressing as blocklin_gemm_kernel (node_kernels.cuh); replaces e3nn o3.Linear
// (sevenn/nn/linear.py:94-100) for self_interaction_1/2 and the self connection, forward and backward.
//
// Arithmetic: error-free tensor-core accumulation.  The tensor core adds into its fp32 accumulator with
// truncation, which biases long fp32/TF32 accumulations (round 1 measured -8.7e-6 eV/atom with 3xTF32).
// Here every operand is first brought to 24-bit fixed point relative to a power-of-two bound of its row
// (A: per (node, component) row over all K, from row_exponent_kernel; W: per output column, on the host)
// and cut into three signed 8-bit slices  q = q0*2^16 + q1*2^8 + q2,  |q_i| <= 128, each exactly
// representable in bf16.  Then
//     a*b*2^-(Ea+Eb-14) = (A0+A1+A2)(B0+B1+B2),   A0 = q0, A1 = q1*2^-8, A2 = q2*2^-16 (same for B)
//   ACC0 = sum_k A0*B0                 integers < 2^23: EXACT in the fp32 accumulator, nothing to truncate
//   ACC1 = sum_k A0*B1+A1*B0+A0*B2+A1*B1+A2*B0     2^-8 of ACC0: its truncation is 2^-32 relative
//   C    = 2^(Ea-7) * 2^(Eb-7) * (ACC0 + ACC1)     one round-to-nearest fp32 add in the epilogue
// (the dropped A1*B2, A2*B1, A2*B2 are < 2^-24 relative).  Six kind::f16 (bf16) MMAs per K = 16 step cost
// the same tensor time as 3xTF32.  Emulated bit for bit on the CPU by tests/test_tc_pack_cpu.py.
//
// Structure (persistent, one CTA per SM, 480 threads = 15 warps, tiles of 128 rows x NT <= 128 columns):
//   warp 12  TMA producer A: per 32-wide K chunk one cp.async.bulk.tensor (3-D map over (k, component,
//            node); 128B swizzle) for the raw fp32 A tile into a 5-deep ring (the HBM/L2 round trip of
//            these loads is what has to be hidden: 80 KB in flight per SM)
//   warp 14  producer W: one cp.async.bulk per chunk for the pre-sliced, pre-arranged W chunk (3-deep ring)
//   warps 0-7 transform: two threads per row of a raw chunk (16 of its 32 k each) pull it into registers
//            (freeing the raw slot), cut it into the three bf16 slices with packed fp32x2 arithmetic and write
//            them in the canonical K-major UMMA layout (8-row x 16-byte core matrices) into a 2-deep operand ring
//   warp 13  TMEM allocation + MMA issuer: one thread issues 12 tcgen05.mma per chunk (M = 128, N = NT, K = 16)
//            into the two TMEM accumulators of the tile; tcgen05.commit frees the operand slots / publishes the tile
//   warps 8-11 epilogue: tcgen05.ld the accumulators (double-buffered in TMEM, so the next tile's MMAs
//            overlap), scale, transpose a 32x32 slab through shared memory (STS.128 / LDS.128) and write C with
//            128-bit row segments, or RED.ADD.F32x4 when accumulating.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "node_kernels.cuh"

namespace s7b {

constexpr int kTcBM = 128;          // rows (nodes) per tile
constexpr int kTcKC = 32;           // K elements per pipeline stage
constexpr int kTcMaxNT = 128;       // columns per tile (two accumulators x two buffers = 512 TMEM columns)
constexpr int kTcThreads = 480;     // 8 transform + 4 epilogue warps, A producer, MMA issuer, W producer
constexpr int kTcXformThreads = 256; // two threads per row of a chunk (16 of its 32 k each)
constexpr int kTcRawBytes = kTcBM * kTcKC * 4;            // 16 KB raw fp32 A chunk
constexpr int kTcASliceBytes = kTcBM * kTcKC * 2;         // 8 KB per bf16 slice
constexpr int kTcBSliceBytes = kTcMaxNT * kTcKC * 2;      // 8 KB per bf16 slice (NT = 128)
// three decoupled rings: raw A chunks (deep: the HBM/L2 round trip of the TMA loads is what has to be
// hidden), sliced A operands and sliced W operands (both released by tcgen05.commit)
constexpr int kTcRawStages = 5, kTcOpsStages = 2, kTcWStages = 3;
constexpr int kTcOpsBytes = 3 * kTcASliceBytes, kTcWBytes = 3 * kTcBSliceBytes;
constexpr int kTcRawOff = 0;
constexpr int kTcOpsOff = kTcRawOff + kTcRawStages * kTcRawBytes;
constexpr int kTcWOff = kTcOpsOff + kTcOpsStages * kTcOpsBytes;
constexpr int kTcEpiOff = kTcWOff + kTcWStages * kTcWBytes;
constexpr int kTcEpiBytes = 4 * 32 * 36 * 4;              // per-warp 32 x 36 transpose scratch (16-byte aligned rows)
constexpr int kTcSmemBytes = kTcEpiOff + kTcEpiBytes + 1024 /*alignment slack*/;
constexpr int kTcZeroRow = -1000;   // row exponent of an all-zero row

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive that carries a REAL data dependency on `dep`: the arrival is predicated on a comparison of the
// value with a bit pattern no fp32 addition can produce, so neither nvcc nor ptxas can drop the dependency
// (a dead `mov` was dropped: the SASS then issued SYNCS.ARRIVE right behind the still outstanding loads and
// the TMA refill overtook them).  The warp therefore waits for the shared-memory loads that produced
// `dep` before it releases the slot they read.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, float dep) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %1, 0x7fc12345;\n"
      "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(__float_as_uint(dep)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA: 3-D tiled tensor load (coordinates innermost first) completing on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// TMA engine, linear form: contiguous global -> shared bulk copy completing on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// lbo = byte distance of core matrices adjacent in K, sbo = of 8-row groups adjacent in M/N.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version for sm_100
  return d;                   // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 2^e as a float (e clamped to the normal range)
__device__ __forceinline__ float exp2i(int e) {
  e = e < -126 ? -126 : (e > 127 ? 127 : e);
  return __int_as_float((e + 127) << 23);
}

// ---- row exponents --------------------------------------------------------------------------------
// E(n, i) with max_k |A[(n,i), k]| < 2^E  (kTcZeroRow for an all-zero / denormal row); one warp per row.
struct RowExpArgs {
  const float* A;
  int* E;                 // [n_nodes, rows_per_node]
  int lda, n_nodes, rows_per_node, nblocks;
  int d[kMaxL], K[kMaxL], a_off[kMaxL], row_base[kMaxL];
};

__global__ void row_exponent_kernel(const RowExpArgs a) {
  // one warp per node: the node's row is read once, coalesced (float4 per lane); every float4 lies inside one
  // (block, component) row (K is a multiple of 4).  Lanes that hold quads of the same row combine their maxima
  // with one warp reduction (match.any + redux.sync), its leader updates the row's slot in shared memory.
  __shared__ unsigned int smax[8][16];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + wib;
  if (lane < 16) smax[wib][lane] = 0u;
  __syncwarp();
  if (n < a.n_nodes) {
    for (int b = 0; b < a.nblocks; ++b) {
      const int K = a.K[b], quads = (a.d[b] * K) >> 2;
      const float4* row = reinterpret_cast<const float4*>(a.A + (size_t)n * a.lda + a.a_off[b]);
      for (int q0 = 0; q0 < quads; q0 += 32) {
        const int q = q0 + lane;
        unsigned int m = 0u;
        int r = -1;
        if (q < quads) {
          const float4 v = __ldg(row + q);
          m = __float_as_uint(v.x) & 0x7fffffffu;
          m = max(m, __float_as_uint(v.y) & 0x7fffffffu);
          m = max(m, __float_as_uint(v.z) & 0x7fffffffu);
          m = max(m, __float_as_uint(v.w) & 0x7fffffffu);
          r = a.row_base[b] + (4 * q) / K;
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, r);
        const unsigned int mm = __reduce_max_sync(peers, m);
        if (r >= 0 && lane == __ffs(peers) - 1) smax[wib][r] = max(smax[wib][r], mm);
        __syncwarp();
      }
    }
  }
  __syncwarp();
  if (n < a.n_nodes && lane < a.rows_per_node) {
    const int ex = (int)(smax[wib][lane] >> 23);
    a.E[(size_t)n * a.rows_per_node + lane] = (ex < 30 || ex == 255) ? kTcZeroRow : ex - 126;     // |a| < 2^(ex-126)
  }
}

// ---- the GEMM --------------------------------------------------------------------------------------
struct TcLinBlock {
  const uint16_t* Wq;   // pre-sliced weights: [nnt][K/32][3 slices][canonical NT x 32 bf16]
  const float* fb;      // [N] column scales 2^(Eb-7)
  int d, K, N, NT, nnt;  // nnt column tiles of NT columns each (nnt * NT >= N; the pad columns carry zero weights)
  int c_off, c_cs;
  int row_base;         // first row of this block in the row-exponent array
  int tile0;            // index of the block's first tile; tiles ordered [node tile][component][n tile]
};
struct TcLinArgs {
  float* C;
  const int* E;         // row exponents of A, [n_nodes, rows_per_node]
  int ldc, n_nodes, rows_per_node, accumulate, nblocks, n_tiles, swizzle;
  int e_bits;           // E holds the raw maxima (|a| bits, written by the producer kernel) instead of exponents
  long long* trace;     // optional timeline of CTA 0 (tools/tc_trace.py): records {role, event, index, clock64}
  int trace_cap;
  TcLinBlock blk[kMaxL];
};
// timeline events of CTA 0 (a.trace != nullptr): five roles, each traced by ONE thread that keeps its own
// record counter in a register and stores {event, index, clock64} with plain stores (no atomics: an atomic's
// round trip would cost more than the stages being measured).  Layout: role r owns records
// [r * trace_cap / 5, (r + 1) * trace_cap / 5); word 0 of the buffer is unused, counts are in words 1..5.
#define TC_TRACE(role, ev, idx)                                                                   \
  do {                                                                                            \
    if (a.trace != nullptr && blockIdx.x == 0) {                                                  \
      const int per = a.trace_cap / 5;                                                            \
      if (trace_n < per) {                                                                        \
        long long* rec = a.trace + 8 + 3 * ((size_t)(role) * per + trace_n);                      \
        rec[0] = (ev);                                                                            \
        rec[1] = (long long)(idx);                                                                \
        rec[2] = clock64();                                                                       \
        ++trace_n;                                                                                \
        a.trace[1 + (role)] = trace_n;                                                            \
      }                                                                                           \
    }                                                                                             \
  } while (0)
struct TcMaps { CUtensorMap m[kMaxL]; };   // per block: A viewed as (k, component, node), fp32

// explicit shared-space accesses (the ring pointers are computed from an aligned base, which makes the compiler
// fall back to generic LD/ST otherwise)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kTcThreads, 1)
blocklin_tc_kernel(const TcLinArgs a, const __grid_constant__ TcMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_raw_full[kTcRawStages], bar_raw_empty[kTcRawStages];
  __shared__ uint64_t bar_ops_full[kTcOpsStages], bar_ops_empty[kTcOpsStages];
  __shared__ uint64_t bar_w_full[kTcWStages], bar_w_empty[kTcWStages];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_sh;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int trace_n = 0;

  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kTcRawStages; ++s) { mbar_init(&bar_raw_full[s], 1); mbar_init(&bar_raw_empty[s], kTcXformThreads); }
    for (int s = 0; s < kTcOpsStages; ++s) { mbar_init(&bar_ops_full[s], kTcXformThreads); mbar_init(&bar_ops_empty[s], 1); }
    for (int s = 0; s < kTcWStages; ++s) { mbar_init(&bar_w_full[s], 1); mbar_init(&bar_w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  // tile t -> (block, node tile, component, n tile)
  auto decode = [&](int t, int& b, int& mt, int& ci, int& nt) {
    b = 0;
    while (b + 1 < a.nblocks && t >= a.blk[b + 1].tile0) ++b;
    const int rel = t - a.blk[b].tile0;
    const int nnt = a.blk[b].nnt;
    nt = rel % nnt;
    const int r2 = rel / nnt;
    ci = r2 % a.blk[b].d;
    mt = r2 / a.blk[b].d;
  };
  auto row_exp = [&](int t, int row) -> int {           // row exponent of tile t's row (kTcZeroRow beyond the nodes)
    if (t >= a.n_tiles) return kTcZeroRow;
    int b, mt, ci, nt;
    decode(t, b, mt, ci, nt);
    const int node = mt * kTcBM + row;
    if (node >= a.n_nodes) return kTcZeroRow;
    int v = __ldg(a.E + (size_t)node * a.rows_per_node + a.blk[b].row_base + ci);
    if (a.e_bits) {
      const int ex = v >> 23;
      v = (ex < 30 || ex == 255) ? kTcZeroRow : ex - 126;          // |a| < 2^(ex-126)
    }
    return v;
  };

  if (warp == 12) {
    // =================== TMA producer, A: raw fp32 chunks ===================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        int b, mt, ci, nt;
        decode(t, b, mt, ci, nt);
        const int n_kc = a.blk[b].K / kTcKC;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % kTcRawStages;
          mbar_wait(&bar_raw_empty[s], ((it / kTcRawStages) & 1) ^ 1);
          TC_TRACE(0, 0, it);      // A producer: slot free, issuing chunk `it`
          mbar_expect_tx(&bar_raw_full[s], (uint32_t)kTcRawBytes);
          tma_load_3d(smem + kTcRawOff + (size_t)s * kTcRawBytes, &maps.m[b], kc * kTcKC, ci, mt * kTcBM, &bar_raw_full[s]);
        }
      }
    }
  } else if (warp == 14) {
    // =================== TMA producer, W: pre-sliced bf16 chunks ===================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        int b, mt, ci, nt;
        decode(t, b, mt, ci, nt);
        const TcLinBlock& B = a.blk[b];
        const int n_kc = B.K / kTcKC;
        const uint32_t b_bytes = 3u * (uint32_t)B.NT * kTcKC * 2u;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(B.Wq) + (size_t)nt * n_kc * b_bytes;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % kTcWStages;
          mbar_wait(&bar_w_empty[s], ((it / kTcWStages) & 1) ^ 1);
          TC_TRACE(4, 0, it);      // W producer: slot free, issuing
          mbar_expect_tx(&bar_w_full[s], b_bytes);
          bulk_load(smem + kTcWOff + (size_t)s * kTcWBytes, wsrc + (size_t)kc * b_bytes, b_bytes, &bar_w_full[s]);
        }
      }
    }
  } else if (warp < 8) {
    // =================== transform: raw fp32 row -> three bf16 slices ===================
    // thread (r, h): row r of the tile, k = 16 h .. 16 h + 15 of the chunk; packed fp32 arithmetic (FFMA2 ...)
    const int r = tid & 127, h = tid >> 7;
    const uint32_t swz = a.swizzle ? (uint32_t)(r & 7) : 0u;
    const uint32_t row_off = (uint32_t)((r & 7) * 16 + (r >> 3) * 512);
    const V2 M2 = splat2(12582912.0f), nM2 = splat2(-12582912.0f);     // 1.5 * 2^23: (x + M) - M = rint(x)
    uint32_t it = 0;
    int Ea_next = row_exp(blockIdx.x, r);
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      int b, mt, ci, nt;
      decode(t, b, mt, ci, nt);
      const int n_kc = a.blk[b].K / kTcKC;
      const int Ea = Ea_next;
      Ea_next = row_exp(t + gridDim.x, r);              // in flight while this tile is converted
      const float sc = (Ea == kTcZeroRow) ? 0.0f : exp2i(23 - Ea);     // t = a * sc, |t| < 2^23
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % kTcRawStages, o = it % kTcOpsStages;
        mbar_wait(&bar_raw_full[s], (it / kTcRawStages) & 1);
        if (tid == 0) TC_TRACE(1, 0, it);    // transform: raw chunk landed
        const uint32_t raw = smem_u32(smem + kTcRawOff + (size_t)s * kTcRawBytes + (size_t)r * 128);
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = lds128(raw + (((uint32_t)(4 * h + q) ^ swz) << 4));
        const float dep = (v[0].x + v[1].x) + (v[2].x + v[3].x);   // touches every load: all four have returned
        mbar_arrive_after(&bar_raw_empty[s], dep);                  // the raw chunk is in registers: its slot can be refilled
        mbar_wait(&bar_ops_empty[o], ((it / kTcOpsStages) & 1) ^ 1);
        if (tid == 0) TC_TRACE(1, 1, it);    // transform: operand slot free
        const uint32_t a0 = smem_u32(smem + kTcOpsOff + (size_t)o * kTcOpsBytes);
        const uint32_t a1 = a0 + kTcASliceBytes, a2 = a1 + kTcASliceBytes;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {                // 8 consecutive k = one 16-byte core-matrix row
          const V2 x[4] = {make_float2(v[2 * kk].x, v[2 * kk].y), make_float2(v[2 * kk].z, v[2 * kk].w),
                           make_float2(v[2 * kk + 1].x, v[2 * kk + 1].y), make_float2(v[2 * kk + 1].z, v[2 * kk + 1].w)};
          uint32_t p0[4], p1[4], p2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const V2 tt = mul_(x[j], sc);
            const V2 q0 = add_(fma_(tt, splat2(1.52587890625e-05f), M2), nM2);        // rint(t / 2^16)
            const V2 r1 = fma_(q0, splat2(-65536.0f), tt);                            // exact
            const V2 q1 = add_(fma_(r1, splat2(0.00390625f), M2), nM2);               // rint(r1 / 2^8)
            const V2 r2 = fma_(q1, splat2(-256.0f), r1);                              // exact
            const V2 q2 = add_(add_(r2, M2), nM2);                                    // rint(r2)
            const V2 s1 = mul_(q1, 0.00390625f), s2 = mul_(q2, 1.52587890625e-05f);
            p0[j] = pack_bf16(q0.x, q0.y);
            p1[j] = pack_bf16(s1.x, s1.y);
            p2[j] = pack_bf16(s2.x, s2.y);
          }
          const uint32_t off = row_off + (uint32_t)(2 * h + kk) * 128u;
          sts128(a0 + off, p0[0], p0[1], p0[2], p0[3]);
          sts128(a1 + off, p1[0], p1[1], p1[2], p1[3]);
          sts128(a2 + off, p2[0], p2[1], p2[2], p2[3]);
        }
        fence_async_smem();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        if (tid == 0) TC_TRACE(1, 2, it);    // transform: slices written
        mbar_arrive(&bar_ops_full[o]);
      }
    }
  } else if (warp == 13) {
    // =================== MMA issuer ===================
    if (lane == 0) {
      uint32_t it = 0, tile_it = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++tile_it) {
        int b, mt, ci, nt;
        decode(t, b, mt, ci, nt);
        const TcLinBlock& B = a.blk[b];
        const int n_kc = B.K / kTcKC;
        const int buf = tile_it & 1;
        mbar_wait(&bar_acc_empty[buf], ((tile_it >> 1) & 1) ^ 1);
        TC_TRACE(2, 2, tile_it);   // MMA: accumulator buffer free
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * 2 * kTcMaxNT)   /* fixed halves: tiles of different NT alternate */;
        const uint32_t acc1 = acc0 + (uint32_t)B.NT;
        const uint32_t idesc = umma_idesc_bf16(B.NT);
        const uint32_t b_slice = (uint32_t)B.NT * kTcKC * 2u;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int o = it % kTcOpsStages, w = it % kTcWStages;
          mbar_wait(&bar_w_full[w], (it / kTcWStages) & 1);
          TC_TRACE(2, 0, it);      // MMA: weights landed
          mbar_wait(&bar_ops_full[o], (it / kTcOpsStages) & 1);
          TC_TRACE(2, 1, it);      // MMA: operands ready, issuing
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + kTcOpsOff + (size_t)o * kTcOpsBytes);
          const uint32_t sb = smem_u32(smem + kTcWOff + (size_t)w * kTcWBytes);
#pragma unroll
          for (int j = 0; j < kTcKC / 16; ++j) {
            const uint32_t ko = (uint32_t)j * 256u;     // two 16-byte K core matrices per MMA
            const uint64_t dA0 = umma_desc(sa + ko, 128, 512);
            const uint64_t dA1 = umma_desc(sa + kTcASliceBytes + ko, 128, 512);
            const uint64_t dA2 = umma_desc(sa + 2 * kTcASliceBytes + ko, 128, 512);
            const uint64_t dB0 = umma_desc(sb + ko, 128, 512);
            const uint64_t dB1 = umma_desc(sb + b_slice + ko, 128, 512);
            const uint64_t dB2 = umma_desc(sb + 2 * b_slice + ko, 128, 512);
            const uint32_t acc = (kc > 0 || j > 0) ? 1u : 0u;
            umma_bf16(acc0, dA0, dB0, idesc, acc);
            umma_bf16(acc1, dA0, dB1, idesc, acc);
            umma_bf16(acc1, dA1, dB0, idesc, 1u);
            umma_bf16(acc1, dA0, dB2, idesc, 1u);
            umma_bf16(acc1, dA1, dB1, idesc, 1u);
            umma_bf16(acc1, dA2, dB0, idesc, 1u);
          }
          umma_commit(&bar_ops_empty[o]);                // both operand slots reusable once these MMAs have read them
          umma_commit(&bar_w_empty[w]);
        }
        umma_commit(&bar_acc_full[buf]);                 // accumulators of this tile complete
      }
    }
  } else if (warp < 12) {
    // =================== epilogue (warps 8..11 <-> TMEM lanes 32*(warp-8) ..) ===================
    const int ew = warp - 8;
    const uint32_t scratch = smem_u32(smem + kTcEpiOff) + (uint32_t)ew * (32 * 36 * 4);
    uint32_t tile_it = 0;
    int Ea_next = row_exp(blockIdx.x, ew * 32 + lane);
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++tile_it) {
      int b, mt, ci, nt;
      decode(t, b, mt, ci, nt);
      const TcLinBlock& B = a.blk[b];
      const int buf = tile_it & 1;
      const int Ea = Ea_next;                                       // the row this thread holds in TMEM
      Ea_next = row_exp(t + gridDim.x, ew * 32 + lane);
      const float fa = (Ea == kTcZeroRow) ? 0.0f : exp2i(Ea - 7);
      const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 2 * kTcMaxNT)   /* fixed halves: tiles of different NT alternate */;
      const int col0 = nt * B.NT;
      const int node0 = mt * kTcBM + ew * 32;
      const int n_rows = min(32, a.n_nodes - node0);                // rows of this warp that exist (<= 0: none)
      float* cbase = a.C + (size_t)node0 * a.ldc + B.c_off + (size_t)ci * B.c_cs + col0;
      // One 32 x 32 slab at a time: thread = TMEM lane = row writes its 32 values as 8 x STS.128 into a
      // [32][36]-float scratch (conflict-free), then the warp re-reads it as 4 rows x 8 column quads per pass
      // (LDS.128) and writes C with 128-bit stores -- or 128-bit fire-and-forget reductions for C += ...:
      // every element of C belongs to exactly one tile, so the single add per element is deterministic and no
      // load latency enters the epilogue.
      const int rsub = lane >> 3, g4 = (lane & 7) * 4;
      bool waited = false;
      for (int c = 0; c < B.NT; c += 32) {
        const int cc = c + g4;                                      // first of this lane's 4 columns
        const bool col_ok = cc < B.NT && col0 + cc < B.N;           // quads are all-in or all-out (NT, N multiples of 4)
        float4 fb4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) fb4 = __ldg(reinterpret_cast<const float4*>(B.fb + col0 + cc));
        if (!waited) {
          mbar_wait(&bar_acc_full[buf], (tile_it >> 1) & 1);
          if (tid == 256) TC_TRACE(3, 0, tile_it);    // epilogue: accumulators complete
          tc_fence_after();
          waited = true;
        }
        {
          uint32_t v0[32], v1[32];
          tmem_ld32(lane_base + (uint32_t)c, v0);
          tmem_ld32(lane_base + (uint32_t)(B.NT + c), v1);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float x0 = (__uint_as_float(v0[4 * q + 0]) + __uint_as_float(v1[4 * q + 0])) * fa;
            const float x1 = (__uint_as_float(v0[4 * q + 1]) + __uint_as_float(v1[4 * q + 1])) * fa;
            const float x2 = (__uint_as_float(v0[4 * q + 2]) + __uint_as_float(v1[4 * q + 2])) * fa;
            const float x3 = (__uint_as_float(v0[4 * q + 3]) + __uint_as_float(v1[4 * q + 3])) * fa;
            sts128(scratch + (uint32_t)(lane * 36 + 4 * q) * 4, __float_as_uint(x0), __float_as_uint(x1), __float_as_uint(x2), __float_as_uint(x3));
          }
        }
        __syncwarp();
        if (col_ok) {
          float4 vals[8];
#pragma unroll
          for (int p8 = 0; p8 < 8; ++p8) vals[p8] = lds128(scratch + (uint32_t)((p8 * 4 + rsub) * 36 + g4) * 4);
          float* p = cbase + (size_t)rsub * a.ldc + cc;
          const size_t step = (size_t)4 * a.ldc;
          if (a.accumulate) {
#pragma unroll
            for (int p8 = 0; p8 < 8; ++p8) {
              if (p8 * 4 + rsub < n_rows)
                atomicAdd(reinterpret_cast<float4*>(p + p8 * step), make_float4(vals[p8].x * fb4.x, vals[p8].y * fb4.y, vals[p8].z * fb4.z, vals[p8].w * fb4.w));
            }
          } else {
#pragma unroll
            for (int p8 = 0; p8 < 8; ++p8) {
              if (p8 * 4 + rsub < n_rows)
                *reinterpret_cast<float4*>(p + p8 * step) = make_float4(vals[p8].x * fb4.x, vals[p8].y * fb4.y, vals[p8].z * fb4.z, vals[p8].w * fb4.w);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (tid == 256) TC_TRACE(3, 1, tile_it);        // epilogue: tile written
      mbar_arrive(&bar_acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace s7b
