// Packed-integer x fp16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), for many tokens
// (the reference's eval loop calls every Linear with M = 2048: opt.py:262-264, llama.py:226-227).
//
//   D[n][m] (TMEM, fp32) = sum_k A[n][k] * B[m][k]
//     A = packed weights, 128 output rows per tile, expanded to fp16 d=(c-cbar)/2^bits by the
//         producer warps and written straight into the 128B-swizzled K-major operand layout
//         (generic-proxy st.shared + fence.proxy.async), never touching HBM as fp16;
//     B = activations x2 (M,K) fp16, BN tokens per tile, staged by TMA (cp.async.bulk.tensor, 128B
//         swizzle, out-of-bounds rows zero-filled so ragged M needs no special case);
//   epilogue: z[m][n] = P_n * D + R_n * xsum[m] (+ bias_n) -> fp16, straight from tcgen05.ld registers.
//
// Warp roles (320 threads, one persistent CTA per SM, static tile schedule):
//   warp 0      TMA producer            warp 1      TMEM alloc + MMA issuer (one lane)
//   warps 2-5   epilogue (TMEM lane quarter = warp % 4)
//   warps 6-9   weight producers: 128-bit loads of packed words -> registers -> fp16 -> smem
// Pipelines: smem ring full[s]/empty[s] (TMA + 4 producer warps -> MMA -> tcgen05.commit), and a
// double-buffered TMEM accumulator tmem_full[a]/tmem_empty[a] (MMA -> epilogue).
#include <cuda.h>

#include "common.cuh"

namespace quip {

constexpr int TC_BM = 128;            // output rows per tile (MMA M)
constexpr int TC_BK = 64;             // k per stage = one 128-byte swizzle atom of fp16
constexpr int TC_THREADS = 320;
constexpr uint32_t TC_WATCHDOG = 1u << 28;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > TC_WATCHDOG) __trap();      // a protocol bug must not hang the GPU
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16, single CTA
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// packed words one producer thread needs for one (row block, lane) over one k super-block (128 k)
template <int BITS>
struct TcWords {
  uint32_t w[BITS == 2 ? 4 : (BITS == 3 ? 6 : 8)];
};
template <int BITS>
__device__ __forceinline__ void tc_load_words(const uint32_t* __restrict__ sb, int l, TcWords<BITS>& r, bool valid) {
  if (!valid) {
#pragma unroll
    for (int i = 0; i < (int)(sizeof(r.w) / 4); ++i) r.w[i] = 0;
    return;
  }
  if constexpr (BITS == 2) {
    uint4 v = *reinterpret_cast<const uint4*>(sb + l * 4);
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  } else if constexpr (BITS == 4) {
    uint4 a = *reinterpret_cast<const uint4*>(sb + l * 4), b = *reinterpret_cast<const uint4*>(sb + 128 + l * 4);
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  } else {
    uint4 a = *reinterpret_cast<const uint4*>(sb + l * 4);
    uint2 b = *reinterpret_cast<const uint2*>(sb + 128 + l * 2);
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    r.w[4] = b.x; r.w[5] = b.y;
  }
}
// expand chunk CH (0..3) of a super-block and store rows g / g+8 of row block `rbl` into the stage's A tile
template <int BITS, int CH>
__device__ __forceinline__ void tc_store_chunk(const TcWords<BITS>& r, uint32_t a_tile, int rbl, int g, int t,
                                               bool valid) {
  uint32_t h[8];
  if (valid) {
    if constexpr (BITS == 2) expand_chunk<2>(r.w[CH], 0u, h);
    else if constexpr (BITS == 4) expand_chunk<4>(r.w[2 * CH], r.w[2 * CH + 1], h);
    else expand_chunk<3, (CH & 1)>(r.w[CH], r.w[4 + (CH >> 1)], h);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0;     // rows beyond N contribute exact zeros
  }
  const int cidx = (CH & 1) * 4 + t;          // 16-byte chunk inside the 128-byte row of this stage
  const uint32_t off = (uint32_t)(rbl * 16 + g) * 128u + (uint32_t)((cidx ^ g) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_tile + off), "r"(h[0]), "r"(h[2]), "r"(h[4]), "r"(h[6])
               : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_tile + off + 8u * 128u), "r"(h[1]), "r"(h[3]),
               "r"(h[5]), "r"(h[7])
               : "memory");
}

template <int BN>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;            // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;    // 256 -> 4, 128 -> 6
  static constexpr int TMEM_COLS = 2 * BN;                     // double-buffered accumulator
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BITS, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
qgemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint32_t* __restrict__ q,
                const float* __restrict__ scales, const float* __restrict__ zeros, const __half* __restrict__ bias,
                const float* __restrict__ xsum, __half* __restrict__ z, int M, int K, int N, int symmetric) {
  using C = TcCfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + (size_t)C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = empty + C::STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + TC_BM - 1) / TC_BM;
  const int tiles_m = (M + BN - 1) / BN;
  const int num_tiles = tiles_n * tiles_m;
  const int KB = K / TC_BK;
  const int KSB = K >> 7;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1 + 4);              // TMA producer + 4 weight-producer warps
      mbar_init(&empty[s], 1);                 // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);            // 4 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], C::B_BYTES);
          tma_load_2d(smem_gen + (size_t)s * C::STAGE_BYTES + C::A_BYTES, &tmap_x, &full[s], kb * TC_BK, m0);
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tmem_empty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_base + (uint32_t)(s * C::STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)     // +32 bytes along K inside the swizzle atom = +2 encoded
            umma_f16_ss(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty[s]);                  // frees the stage once these MMAs have read it
          if (kb == KB - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++s == C::STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ================= epilogue: TMEM -> registers -> z =================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int n = (tile % tiles_n) * TC_BM + quarter * 32 + lane;
      const int m0 = (tile / tiles_n) * BN;
      float Pn = 0.f, Rn = 0.f, bn = 0.f;
      if (n < N) {
        float sc = scales[n];
        Pn = sc * (float)(1 << BITS);
        if (!symmetric) Rn = sc * (0.5f * (float)((1 << BITS) - 1)) - zeros[n];
        if (bias) bn = __half2float(bias[n]);
      }
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int m = m0 + c0 + c;
            if (m < M) {
              float v = Pn * __uint_as_float(r[c]) + bn;
              if (!symmetric) v += Rn * __ldg(&xsum[m]);
              z[(int64_t)m * N + n] = __float2half_rn(v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  } else {
    // ================= weight producers: packed words -> fp16 operand tile =================
    const int pw = warp - 6;                       // 0..3
    const int g = lane & 7, t = lane >> 3;         // 8 consecutive lanes = 8 rows g: conflict-free st.shared.v4
    const int l = 4 * g + t;                       // "lane" index of the native layout
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int rb_base = (tile % tiles_n) * (TC_BM / 16);
      const int NRB = N >> 4;
      // this thread serves row blocks pw and pw+4 of the tile
      const int rbl0 = pw, rbl1 = pw + 4;
      const bool v0 = rb_base + rbl0 < NRB, v1 = rb_base + rbl1 < NRB;
      const uint32_t* q0 = q + (int64_t)(rb_base + rbl0) * KSB * sb_words(BITS);
      const uint32_t* q1 = q + (int64_t)(rb_base + rbl1) * KSB * sb_words(BITS);
      TcWords<BITS> c0, c1, n0, n1;                // current and next super-block (software prefetch)
      tc_load_words<BITS>(q0, l, c0, v0);
      tc_load_words<BITS>(q1, l, c1, v1);
      for (int ksb = 0; ksb < KSB; ++ksb) {
        const bool more = ksb + 1 < KSB;
        tc_load_words<BITS>(q0 + (int64_t)(ksb + 1) * sb_words(BITS), l, n0, v0 && more);
        tc_load_words<BITS>(q1 + (int64_t)(ksb + 1) * sb_words(BITS), l, n1, v1 && more);
#pragma unroll
        for (int half = 0; half < 2; ++half) {     // two 64-k stages per 128-k super-block
          mbar_wait(&empty[s], ph ^ 1u);
          const uint32_t a_tile = smem_base + (uint32_t)(s * C::STAGE_BYTES);
          if (half == 0) {
            tc_store_chunk<BITS, 0>(c0, a_tile, rbl0, g, t, v0);
            tc_store_chunk<BITS, 1>(c0, a_tile, rbl0, g, t, v0);
            tc_store_chunk<BITS, 0>(c1, a_tile, rbl1, g, t, v1);
            tc_store_chunk<BITS, 1>(c1, a_tile, rbl1, g, t, v1);
          } else {
            tc_store_chunk<BITS, 2>(c0, a_tile, rbl0, g, t, v0);
            tc_store_chunk<BITS, 3>(c0, a_tile, rbl0, g, t, v0);
            tc_store_chunk<BITS, 2>(c1, a_tile, rbl1, g, t, v1);
            tc_store_chunk<BITS, 3>(c1, a_tile, rbl1, g, t, v1);
          }
          fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
        c0 = n0;
        c1 = n1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static int g_num_sms = 0;

template <int BITS, int BN>
static int launch_tc(const QuipLinearDesc* d, const __half* x, const float* xsum, const __half* bias, __half* z,
                     int M, cudaStream_t s) {
  using C = TcCfg<BN>;
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return QUIP_ERR_CUDA;
  }
  CUtensorMap tmap;
  cuuint64_t dims[2] = {(cuuint64_t)d->K, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)d->K * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)x, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (x=%p M=%d K=%d)", (int)r, (const void*)x, M, d->K);
    return QUIP_ERR_CUDA;
  }
  if (!g_num_sms) {
    int dev = 0;
    QUIP_CUDA(cudaGetDevice(&dev));
    QUIP_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  auto kern = qgemm_tc_kernel<BITS, BN>;
  QUIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  int tiles = ceil_div(d->N, TC_BM) * ceil_div(M, BN);
  int grid = tiles < g_num_sms ? tiles : g_num_sms;
  kern<<<grid, TC_THREADS, C::SMEM, s>>>(tmap, reinterpret_cast<const uint32_t*>(d->qweight), d->scales, d->zeros, bias,
                                         xsum, z, M, d->K, d->N, (d->flags & QUIP_FLAG_SYMMETRIC) ? 1 : 0);
  QUIP_LAUNCHED("qgemm_tc_kernel");
  return QUIP_OK;
}

int qgemm_tc(const QuipLinearDesc* d, const __half* x, const float* xsum, const __half* bias, __half* z, int M,
             cudaStream_t s) {
  QUIP_CHECK_ARG(((uintptr_t)x & 15) == 0, "tcgen05 path needs 16-byte aligned activations");
  const bool wide = M > 128;
#define QUIP_TC(B)                                                            \
  if (d->bits == B)                                                           \
    return wide ? launch_tc<B, 256>(d, x, xsum, bias, z, M, s) : launch_tc<B, 128>(d, x, xsum, bias, z, M, s);
  QUIP_TC(2) QUIP_TC(3) QUIP_TC(4)
#undef QUIP_TC
  set_error("tcgen05 kernel: unsupported bits=%d", d->bits);
  return QUIP_ERR_UNSUPPORTED;
}

}  // namespace quip
