// 2-CTA (cta_group::2) variant of the packed tcgen05 GEMM: a cluster of two CTAs on one TPC computes a
// 256 (output rows) x 256 (tokens) tile.  Each CTA expands the packed words of its own 128 rows and
// TMA-loads only HALF of the activation tile (128 tokens); one tcgen05.mma.cta_group::2 issued by the
// leader CTA reads A from both CTAs' shared memory (M = 256) and the two activation halves as one N = 256
// operand.  Per SM this halves the activation bytes pulled from L2 and written to / read from shared
// memory, which is what bounds the 1-CTA kernel (qgemm_tc.cu) at ~60-65 % of the tensor peak.
//
// Barrier protocol (s = smem stage, a = accumulator buffer):
//   full[s]       leader CTA only; 1 (leader TMA thread, expect_tx = both halves) + 4 + 4 producer warps
//                 (the peer's arrive remotely through mapa); both CTAs' TMA complete_tx on it
//   empty[s]      in both CTAs; tcgen05.commit multicast from the leader frees the stage in both
//   tmem_full[a]  in both CTAs; commit multicast
//   tmem_empty[a] leader only; 4 + 4 epilogue warps (peer's arrive remotely)
// All waits are plain CTA-scope try_waits on barriers in the waiter's own shared memory (as CUTLASS' 2-SM
// kernels do): a cluster-scope acquire compiles to MEMBAR.ALL.GPU + CCTL.IVALL after every wait, which
// drains the producers' in-flight prefetches and the epilogue's stores (measured: 2x slower).
#include "tc_common.cuh"

namespace quip {

struct T2Cfg {
  static constexpr int BN = 256;                       // tokens per cluster tile
  static constexpr int HALF = 128;                     // tokens staged per CTA
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;    // 16 KB
  static constexpr int B_BYTES = HALF * TC_BK * 2;     // 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int TMEM_COLS = 512;                // two 256-column accumulators
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + 8192;
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
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

template <int BITS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
qgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint32_t* __restrict__ q,
                 const float* __restrict__ scales, const float* __restrict__ zeros, const __half* __restrict__ bias,
                 const float* __restrict__ xsum, __half* __restrict__ z, int M, int K, int N, int symmetric) {
  using C = T2Cfg;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + (size_t)C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  (used in the leader)
  uint64_t* empty = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = empty + C::STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]       (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  __half* epi_stage = reinterpret_cast<__half*>(smem_gen + (size_t)C::STAGES * C::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tiles_n = (N + 2 * TC_BM - 1) / (2 * TC_BM);
  const int tiles_m = (M + C::BN - 1) / C::BN;
  const int num_tiles = tiles_n * tiles_m;
  const int KB = K / TC_BK;
  const int KSB = K >> 7;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1 + 4 + 4);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4 + 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                          // barriers of both CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: this CTA's half of the activation tile =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / tiles_n) * C::BN + (int)rank * C::HALF;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          const uint32_t leader_full = mapa_u32(smem_u32(&full[s]), 0);
          if (leader) mbar_arrive_expect_tx(&full[s], 2 * C::B_BYTES);
          tma_load_2d_2sm(smem_base + (uint32_t)(s * C::STAGE_BYTES + C::A_BYTES), &tmap_x, leader_full, kb * TC_BK, m0);
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(C::BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        if (lane == 0) mbar_wait(&tmem_empty[as], aph ^ 1u);
        __syncwarp();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * C::BN);
        for (int kb = 0; kb < KB; ++kb) {
          if (lane == 0) mbar_wait(&full[s], ph);
          __syncwarp();
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_base + (uint32_t)(s * C::STAGE_BYTES);
            const uint64_t adesc = make_sw128_desc(a_addr);
            const uint64_t bdesc = make_sw128_desc(a_addr + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16_ss_2cta(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit_2cta(&empty[s]);
            if (kb == KB - 1) umma_commit_2cta(&tmem_full[as]);
          }
          __syncwarp();
          if (++s == C::STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < 6) {
    // ================= epilogue: this CTA's 128 rows x 256 tokens =================
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int n = (tile % tiles_n) * 2 * TC_BM + (int)rank * TC_BM + quarter * 32 + lane;
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
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[as]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
      }
    }
  } else {
    // ================= weight producers: this CTA's 128 rows =================
    const int pw = (warp - 6) & 3, grp = (warp - 6) >> 2;
    const uint32_t leader_full0 = mapa_u32(smem_u32(&full[0]), 0);
    weight_producer_loop<BITS, C::STAGES, C::STAGE_BYTES>(
        pw, grp, lane, q, KSB, N, empty, smem_base, cluster_id, num_clusters, num_tiles,
        [&](int tile) { return (tile % tiles_n) * (2 * TC_BM / 16) + (int)rank * (TC_BM / 16); },
        [&](int s) {
          if (leader) mbar_arrive(&full[s]);
          else mbar_arrive_cluster(leader_full0 + (uint32_t)(s * 8));
        });
  }

  tc_fence_before();
  cluster_sync_all();                          // the peer's shared memory / TMEM stay alive until every MMA retired
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

int make_act_map(CUtensorMap* tmap, const void* x, int64_t rows, int64_t cols, int box_rows);
int num_sms();

template <int BITS>
static int launch_tc2(const QuipLinearDesc* d, const __half* x, const float* xsum, const __half* bias, __half* z, int M,
                      cudaStream_t s) {
  using C = T2Cfg;
  CUtensorMap tmap;
  if (int e = make_act_map(&tmap, x, M, d->K, C::HALF)) return e;
  auto kern = qgemm_tc2_kernel<BITS>;
  QUIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  int tiles = ceil_div(d->N, 2 * TC_BM) * ceil_div(M, C::BN);
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  kern<<<2 * clusters, TC_THREADS, C::SMEM, s>>>(tmap, reinterpret_cast<const uint32_t*>(d->qweight), d->scales, d->zeros,
                                                 bias, xsum, z, M, d->K, d->N, (d->flags & QUIP_FLAG_SYMMETRIC) ? 1 : 0);
  QUIP_LAUNCHED("qgemm_tc2_kernel");
  return QUIP_OK;
}

int qgemm_tc2(const QuipLinearDesc* d, const __half* x, const float* xsum, const __half* bias, __half* z, int M,
              cudaStream_t s) {
  QUIP_CHECK_ARG(((uintptr_t)x & 15) == 0, "tcgen05 path needs 16-byte aligned activations");
  if (d->bits == 2) return launch_tc2<2>(d, x, xsum, bias, z, M, s);
  if (d->bits == 3) return launch_tc2<3>(d, x, xsum, bias, z, M, s);
  if (d->bits == 4) return launch_tc2<4>(d, x, xsum, bias, z, M, s);
  set_error("tcgen05 2-CTA kernel: unsupported bits=%d", d->bits);
  return QUIP_ERR_UNSUPPORTED;
}

}  // namespace quip
