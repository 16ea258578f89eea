// TS-mode packed tcgen05 GEMM: the expanded weights never touch shared memory.
//
// In the SS kernel (qgemm_tc.cu) every 64-k stage moves 48 KB of MMA operand reads + 32 KB of TMA writes +
// 16 KB of expanded weights through the 128 B/cycle shared-memory pipe -- measured 961 wavefronts per stage
// against a 512-cycle MMA budget, i.e. the kernel is shared-memory bound at ~55-65 % of the tensor peak.
// Here the A operand (weights, 128 rows x 64 k per stage) is written by the producer warps straight from
// registers into TENSOR MEMORY with tcgen05.st (lane = output row, 32-bit column = two consecutive k) and the
// MMA is issued in its TS form (A from TMEM, B = activations from shared memory).  Shared memory then only
// carries the activation tile: TMA write + one MMA read.
//
// TMEM budget (512 columns): two 192-column fp32 accumulators (BN = 192 tokens per tile, double buffered so
// the epilogue overlaps the next tile) + a 4-slot ring of 32-column A stages.
//
// The native packed layout gives lane (g,t) of a warp the codes of rows {g, g+8} x 8 k; tcgen05.st.32x32b wants
// thread i to own row i of its warp's 32-lane quarter.  The 8 packed words a row needs per stage are fetched
// from the 4 lanes that hold them with warp shuffles (16 SHFL per stage), then expanded locally.
#include "tc_common.cuh"

namespace quip {

struct TsCfg {
  static constexpr int BN = 192;
  static constexpr int STAGES = 4;
  static constexpr int B_BYTES = BN * TC_BK * 2;            // 24 KB
  static constexpr int A_COLS = TC_BK / 2;                  // 32 TMEM columns per stage
  static constexpr int A_COL0 = 2 * BN;                     // A ring starts after the two accumulators
  static constexpr int TMEM_COLS = 512;
  static constexpr size_t SMEM = (size_t)STAGES * B_BYTES + 1024 + 256 + 8192;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 2-bit only (the headline width); 3- and 4-bit layers use the SS kernel.
__global__ void __launch_bounds__(TC_THREADS, 1)
qgemm_ts_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint32_t* __restrict__ q,
                const float* __restrict__ scales, const float* __restrict__ zeros, const __half* __restrict__ bias,
                const float* __restrict__ xsum, __half* __restrict__ z, int M, int K, int N, int symmetric) {
  using C = TsCfg;
  constexpr int BITS = 2;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + (size_t)C::STAGES * C::B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  __half* epi_stage = reinterpret_cast<__half*>(smem_gen + (size_t)C::STAGES * C::B_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + TC_BM - 1) / TC_BM;
  const int tiles_m = (M + C::BN - 1) / C::BN;
  const int num_tiles = tiles_n * tiles_m;
  const int KB = K / TC_BK;
  const int KSB = K >> 7;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1 + 4);              // TMA producer + the 4 producer warps of one group
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
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
        const int m0 = (tile / tiles_n) * C::BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], C::B_BYTES);
          tma_load_2d(smem_gen + (size_t)s * C::B_BYTES, &tmap_x, &full[s], kb * TC_BK, m0);
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (A from TMEM) =================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(C::BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tmem_empty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * C::BN);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t bdesc = make_sw128_desc(smem_base + (uint32_t)(s * C::B_BYTES));
          const uint32_t tmem_a = tmem_base + (uint32_t)(C::A_COL0 + s * C::A_COLS);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_f16_ts(tmem_d, tmem_a + (uint32_t)(8 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == KB - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++s == C::STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ================= epilogue =================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int n = (tile % tiles_n) * TC_BM + quarter * 32 + lane;
      const int m0 = (tile / tiles_n) * C::BN;
      float Pn = 0.f, Rn = 0.f, bn = 0.f;
      if (n < N) {
        float sc = scales[n];
        Pn = sc * (float)(1 << BITS);
        if (!symmetric) Rn = sc * (0.5f * (float)((1 << BITS) - 1)) - zeros[n];
        if (bias) bn = __half2float(bias[n]);
      }
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * C::BN);
      __half* stage = epi_stage + (warp - 2) * EPI_STAGE_HALVES;
      const int n_warp = n - lane;
#pragma unroll 1
      for (int c0 = 0; c0 < C::BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          v[c] = Pn * __uint_as_float(r[c]) + bn;
          if (!symmetric) {
            const int m = m0 + c0 + c;
            v[c] += Rn * (m < M ? __ldg(&xsum[m]) : 0.f);
          }
        }
        epilogue_store_chunk(stage, v, z, N, m0 + c0, M, n_warp, N, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  } else {
    // ================= weight producers: packed words -> registers -> TMEM =================
    const int quarter = warp & 3;                  // TMEM lanes this warp may touch: 32*quarter ..
    const int grp = (warp - 6) >> 2;
    const int r16 = lane & 15, g = r16 & 7, hi = r16 >> 3, use_b = lane >> 4;
    const int NRB = N >> 4;
    uint32_t it_base = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it_base += 2u * (uint32_t)KSB) {
      const int rb_a = (tile % tiles_n) * (TC_BM / 16) + 2 * quarter, rb_b = rb_a + 1;
      const bool va = rb_a < NRB, vb = rb_b < NRB;
      const uint32_t* qa = q + (int64_t)(va ? rb_a : 0) * KSB * sb_words(BITS);
      const uint32_t* qb = q + (int64_t)(vb ? rb_b : 0) * KSB * sb_words(BITS);
      uint4 ca = make_uint4(0, 0, 0, 0), cb = ca, na = ca, nb = ca;
      if (grp < KSB) {
        if (va) ca = *reinterpret_cast<const uint4*>(qa + (int64_t)grp * sb_words(BITS) + lane * 4);
        if (vb) cb = *reinterpret_cast<const uint4*>(qb + (int64_t)grp * sb_words(BITS) + lane * 4);
      }
      for (int ksb = grp; ksb < KSB; ksb += TC_PROD_GROUPS) {
        if (ksb + TC_PROD_GROUPS < KSB) {
          if (va) na = *reinterpret_cast<const uint4*>(qa + (int64_t)(ksb + TC_PROD_GROUPS) * sb_words(BITS) + lane * 4);
          if (vb) nb = *reinterpret_cast<const uint4*>(qb + (int64_t)(ksb + TC_PROD_GROUPS) * sb_words(BITS) + lane * 4);
        }
        const uint32_t wa[4] = {ca.x, ca.y, ca.z, ca.w}, wb[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t regs[32];
#pragma unroll
          for (int cl = 0; cl < 2; ++cl)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t xa = __shfl_sync(0xffffffffu, wa[2 * half + cl], 4 * g + t);
              const uint32_t xb = __shfl_sync(0xffffffffu, wb[2 * half + cl], 4 * g + t);
              const uint32_t w = (use_b ? xb : xa) >> (2 * hi);      // this row's pairs now sit at even slots
              regs[cl * 16 + t * 4 + 0] = dq2<slot2(0, 0)>(w);
              regs[cl * 16 + t * 4 + 1] = dq2<slot2(1, 0)>(w);
              regs[cl * 16 + t * 4 + 2] = dq2<slot2(2, 0)>(w);
              regs[cl * 16 + t * 4 + 3] = dq2<slot2(3, 0)>(w);
            }
          if (!(use_b ? vb : va)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) regs[i] = 0;               // rows beyond N contribute exact zeros
          }
          const uint32_t it = it_base + 2u * (uint32_t)ksb + (uint32_t)half;
          const int s = (int)(it % (uint32_t)C::STAGES);
          const uint32_t ph = (it / (uint32_t)C::STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          tc_fence_after();
          tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(C::A_COL0 + s * C::A_COLS), regs);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
        }
        ca = na;
        cb = nb;
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

int make_act_map(CUtensorMap* tmap, const void* x, int64_t rows, int64_t cols, int box_rows);
int num_sms();

int qgemm_ts(const QuipLinearDesc* d, const __half* x, const float* xsum, const __half* bias, __half* z, int M,
             cudaStream_t s) {
  using C = TsCfg;
  QUIP_CHECK_ARG(d->bits == 2, "TS-mode kernel is built for 2-bit weights (got %d)", d->bits);
  QUIP_CHECK_ARG(((uintptr_t)x & 15) == 0, "tcgen05 path needs 16-byte aligned activations");
  CUtensorMap tmap;
  if (int e = make_act_map(&tmap, x, M, d->K, C::BN)) return e;
  QUIP_CUDA(cudaFuncSetAttribute(qgemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  int tiles = ceil_div(d->N, TC_BM) * ceil_div(M, C::BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  qgemm_ts_kernel<<<grid, TC_THREADS, C::SMEM, s>>>(tmap, reinterpret_cast<const uint32_t*>(d->qweight), d->scales,
                                                   d->zeros, bias, xsum, z, M, d->K, d->N,
                                                   (d->flags & QUIP_FLAG_SYMMETRIC) ? 1 : 0);
  QUIP_LAUNCHED("qgemm_ts_kernel");
  return QUIP_OK;
}

}  // namespace quip
