// Block-diagonal irreps linear on the 5th-generation tensor cores (tcgen05, sm_100a):
//   C[(n,i), :N] (+)= A[(n,i), :K] * W[K, N]        fp32 in / fp32 out, 3xTF32 inside
// Same operand addressing as blocklin_gemm_kernel (node_kernels.cuh); replaces e3nn o3.Linear
// (sevenn/nn/linear.py:94-100) for self_interaction_1/2 and the self connection.
//
// Precision: every fp32 operand is split a = a_hi + a_lo with a_hi = rna_tf32(a) and
// a_lo = rna_tf32(a - a_hi) (round-to-nearest, so the hardware's truncation to TF32 loses nothing and
// the residual 2^-23-level errors are unbiased); the tensor core accumulates all four products
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo + a_lo*b_lo in fp32 in TMEM.  Weights are pre-split once.
//
// Structure (one CTA = 128 threads = one 128-row tile x one N chunk of <= 256 columns):
//   * all threads stage a [128 x 32] A chunk (split hi/lo on the fly) and the matching
//     [N x 32] W^T chunk into shared memory in the canonical K-major, no-swizzle UMMA layout
//     (8-row x 16-byte core matrices; LBO = 128 B along K, SBO = 1024 B between row groups);
//   * one elected thread issues 16 tcgen05.mma.kind::tf32 (M=128, N, K=8) per chunk and commits to an
//     mbarrier; two stages, so the loads of chunk c+1 overlap the MMAs of chunk c;
//   * the accumulator [128 x N] fp32 lives in TMEM; after the last commit each warp reads its 32
//     lanes with tcgen05.ld.32x32b.x32 and writes the rows out (optionally C += ...).
#pragma once
#include "common.cuh"
#include "node_kernels.cuh"

namespace s7b {

constexpr int kTcBM = 128, kTcKC = 32, kTcThreads = 128, kTcMaxN = 256;

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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version for sm_100
  return d;                   // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// round-to-nearest conversion to a TF32-representable fp32 value (the tensor core itself would truncate)
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// canonical K-major no-swizzle offset (bytes) of the 16-byte chunk (row r, k-quad q) in a [rows x 32] tile
__device__ __forceinline__ uint32_t canon_off(int r, int q) { return (uint32_t)((r & 7) * 16 + (r >> 3) * 1024 + q * 128); }

struct TcLinBlock {
  const float* Wt_hi;   // [N, K] row-major (= W^T), truncated to TF32
  const float* Wt_lo;   // [N, K] remainder
  int d, K, N;
  int a_off, a_cs, c_off, c_cs;
};
struct TcLinArgs {
  const float* A;
  float* C;
  int lda, ldc, n_nodes, accumulate, nblocks;
  TcLinBlock blk[kMaxL];
};

// dynamic smem: 2 stages x (A_hi 16K + A_lo 16K + B_hi nb*128 + B_lo nb*128) ; grid = (row tiles, N chunks, blocks)
__global__ void __launch_bounds__(kTcThreads, 1) blocklin_tc_kernel(const TcLinArgs a, int n_chunk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_sh;

  const TcLinBlock b = a.blk[blockIdx.z];
  const int rows = a.n_nodes * b.d;
  const int row0 = blockIdx.x * kTcBM;
  const int col0 = blockIdx.y * n_chunk;
  if (row0 >= rows || col0 >= b.N) return;        // uniform per CTA
  const int nb = min(n_chunk, b.N - col0);         // columns of this CTA (multiple of 16)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t stage_bytes = 2u * 16384u + 2u * (uint32_t)n_chunk * 128u;
  uint8_t* stage_ptr[2] = {smem, smem + stage_bytes};

  // TMEM columns: power of two >= 32 covering nb
  uint32_t ncols = 32;
  while ((int)ncols < nb) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  // A row of this thread
  const int gr = row0 + tid;
  const float* a_row = nullptr;
  if (gr < rows) {
    const int n = gr / b.d, i = gr - n * b.d;
    a_row = a.A + (size_t)n * a.lda + b.a_off + i * b.a_cs;
  }
  const uint32_t idesc = umma_idesc_tf32(nb);
  const int n_kc = b.K / kTcKC;

  for (int kc = 0; kc < n_kc; ++kc) {
    const int s = kc & 1;
    if (kc >= 2) mbar_wait(&mma_done[s], (uint32_t)(((kc >> 1) - 1) & 1));   // stage free again
    uint8_t* sA_hi = stage_ptr[s];
    uint8_t* sA_lo = sA_hi + 16384;
    uint8_t* sB_hi = sA_lo + 16384;
    uint8_t* sB_lo = sB_hi + (size_t)n_chunk * 128;
    // ---- A chunk: row tid, k = kc*32 .. +31, split hi / lo
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_row != nullptr) v = __ldg(reinterpret_cast<const float4*>(a_row + kc * kTcKC + 4 * q));
      float4 h, l;
      h.x = rna_tf32(v.x); l.x = rna_tf32(v.x - h.x);
      h.y = rna_tf32(v.y); l.y = rna_tf32(v.y - h.y);
      h.z = rna_tf32(v.z); l.z = rna_tf32(v.z - h.z);
      h.w = rna_tf32(v.w); l.w = rna_tf32(v.w - h.w);
      const uint32_t off = canon_off(tid, q);
      *reinterpret_cast<float4*>(sA_hi + off) = h;
      *reinterpret_cast<float4*>(sA_lo + off) = l;
    }
    // ---- W^T chunk: rows col0 .. col0+nb-1
    for (int r = tid; r < nb; r += kTcThreads) {
      const float* wh = b.Wt_hi + (size_t)(col0 + r) * b.K + kc * kTcKC;
      const float* wl = b.Wt_lo + (size_t)(col0 + r) * b.K + kc * kTcKC;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t off = canon_off(r, q);
        *reinterpret_cast<float4*>(sB_hi + off) = __ldg(reinterpret_cast<const float4*>(wh + 4 * q));
        *reinterpret_cast<float4*>(sB_lo + off) = __ldg(reinterpret_cast<const float4*>(wl + 4 * q));
      }
    }
    fence_async_smem();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aH = smem_u32(sA_hi), aL = smem_u32(sA_lo), bH = smem_u32(sB_hi), bL = smem_u32(sB_lo);
#pragma unroll
      for (int j = 0; j < kTcKC / 8; ++j) {
        const uint32_t ko = (uint32_t)j * 256u;   // two 16-byte K chunks per MMA
        const uint64_t dAh = umma_desc(aH + ko, 128, 1024), dAl = umma_desc(aL + ko, 128, 1024);
        const uint64_t dBh = umma_desc(bH + ko, 128, 1024), dBl = umma_desc(bL + ko, 128, 1024);
        umma_tf32(tmem_d, dAh, dBh, idesc, (kc > 0 || j > 0) ? 1u : 0u);
        umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
        umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
        umma_tf32(tmem_d, dAl, dBl, idesc, 1u);
      }
      umma_commit(&mma_done[s]);
    }
  }
  // ---- wait for the last commit (covers every MMA issued before it)
  {
    const int last = n_kc - 1;
    mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
  }
  tc_fence_after();

  // ---- epilogue: warp w owns TMEM lanes 32w .. 32w+31 == rows row0 + 32w + lane
  const int r = row0 + warp * 32 + lane;
  float* c_row = nullptr;
  if (r < rows) {
    const int n = r / b.d, ii = r - n * b.d;
    c_row = a.C + (size_t)n * a.ldc + b.c_off + ii * b.c_cs + col0;
  }
  for (int c = 0; c < nb; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    if (c_row != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (c + j < nb) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          float4* p = reinterpret_cast<float4*>(c_row + c + j);
          if (a.accumulate) {
            const float4 old = *p;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *p = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
  }
}

}  // namespace s7b
