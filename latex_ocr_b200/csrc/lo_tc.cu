// tcgen05 + TMA kernels (sm_100a): NHWC 3x3 implicit-GEMM convolution and the K-major "NT" GEMM.
//
//   D[128 x NT] (TMEM, fp32) += A[128 x 64] (smem, bf16, K-major, SWIZZLE_128B) * B[NT x 64]^T (smem, bf16, K-major)
//
// A tiles come from TMA: for the convolution a 4-D tiled map over the NHWC feature map {C, W, H, N} with box
// {64, BW, BH, 1} (BW*BH = 128 output positions); tap (r,s) is just a coordinate shift and the halo / zero padding
// falls out of TMA's out-of-bounds zero fill — no im2col buffer, no predicates.  B tiles are rows of the
// [Cout][9*Cin] weight matrix.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU/mask -> global).  3-4 stage mbarrier ring, one output
// tile per CTA, 2 CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "lo_common.cuh"
#include "lo_ptx.cuh"

namespace lo {

// ------------------------------------------------------------------------------------------------
// PTX wrappers (strings follow cute/arch/{copy_sm90_tma,mma_sm100_umma,tmem_allocator_sm100}.hpp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// multicast variants: the box lands at the same CTA-relative offset in every CTA of `mask` and completes on each one's mbarrier
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6, %7}], [%2], %3;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (quadrant base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 B, 8-row groups
// 1024 B apart (SBO), LBO unused for swizzled K-major (=1), version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address   bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (16 B units) bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset  bits [32,46)
  d |= (uint64_t)1 << 46;                           // version = 1         bits [46,48)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B        bits [61,64)
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
  // problem
  int M, N, K;            // GEMM view: M rows (positions), N = Cout, K = 9*Cin (conv) or K (gemm)
  int conv;               // 0: plain NT GEMM ; 1: 3x3 conv
  int Ho, Wo, Cin, pad;   // conv geometry (output H, W)
  int BW, BH, tiles_w, tiles_h;
  // epilogue
  const float* bias;
  const bf16* mask;
  void* out;
  int64_t ldc;
  int out_f32, accumulate, relu;
  int kb_per_split, atomic;   // split-K over gridDim.z: fp32 atomics onto `out` (bias added by split 0)
  int w_evict_last;           // keep the B (weight) tiles in L2: the per-step decoder GEMMs re-read them every step
  int half_w, half_h;         // MC=1 conv: box offset of the second half of the A tile (one of them is 0)
  long long* dbg;             // optional: clock64 stamps of CTA (0,0,0) at the pipeline milestones (lo_debug_buffer)
  // fused LSTM-cell epilogue (decoder forward): the GEMM's N dimension is gate-interleaved (column 4*j + gate), the
  // epilogue adds the embedding-table row and the recurrent projection, applies the cell and writes h, c, gates
  int lstm;
  const float* l_ptab; const int64_t* l_tok; int64_t l_tok_stride;
  const float* l_hh; int64_t l_hh_stride;
  const float* l_cprev; float* l_gates; float* l_c; float* l_h; bf16* l_hbf;
  float* l_hd; int64_t l_hd_stride; const float* l_dmask;
  int l_D, l_V;
};

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_THREADS = 224;   // warp 0: A-tile TMA producer, 1: TMEM + MMA issuer, 2..5: epilogue, 6: B-tile TMA producer

template <int NT, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;      // 16 KB
  static constexpr int B_BYTES = NT * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// MC = 1: the two CTAs of a (2,1,1) cluster compute neighbouring N tiles of the SAME 128-row A tile; each loads one
// 64-row half of A and multicasts it to both (halves the L2 -> SM traffic of A; the kernel is L2-bandwidth bound at
// 128x128 tiles).  mapA then describes HALF boxes.
template <int NT, int STAGES, int MC>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_conv_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, TcParams p) {
  using SM = TcSmem<NT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  __shared__ __align__(16) float s_bias[NT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NT;
  const int mt = blockIdx.y;
  const int KB_all = p.K / TC_BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int KB = min(KB_all, kb0 + p.kb_per_split) - kb0;      // K blocks of this split

  // tile origin
  int img = 0, h0 = 0, w0 = 0, m0 = mt * TC_BM;
  if (p.conv) {
    const int per_img = p.tiles_w * p.tiles_h;
    img = mt / per_img;
    const int rem = mt % per_img;
    h0 = (rem / p.tiles_w) * p.BH;
    w0 = (rem % p.tiles_w) * p.BW;
  }

  const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar + s, 2);                 // two producer threads (A tiles / B tiles): one TMA issue costs ~200 cycles,
                                                  // a single producer (~410 cycles per K block) could not feed the 256-cycle MMAs
      mbar_init(empty_bar + s, MC ? 2 : 1);       // MC: the slot is also written by the peer -> both MMA threads release it
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NT);
  pdl_wait();          // everything above touched only shared memory / TMEM / kernel parameters
  pdl_trigger();
  tc_fence_before();
  if (MC) cluster_sync_all(); else __syncthreads();   // peer barriers must be initialised before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = MC ? cluster_ctarank() : 0;
  if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer, A tiles =====
      const int cpb = p.conv ? p.Cin / TC_BK : 1;
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        uint8_t* sa = smem + s * SM::STAGE_BYTES;
        mbar_expect_tx(full_bar + s, SM::A_BYTES);
        const int kg = kb0 + kb;
        if (MC) {
          uint8_t* sh = sa + crank * (SM::A_BYTES / 2);        // my half of the tile, delivered to both CTAs
          if (p.conv) {
            const int tap = kg / cpb, cb = kg % cpb;
            const int r = tap / 3, q = tap % 3;
            tma_load_4d_mc(sh, &mapA, full_bar + s, cb * TC_BK, w0 + (int)crank * p.half_w + q - p.pad,
                           h0 + (int)crank * p.half_h + r - p.pad, img, (uint16_t)3);
          } else {
            tma_load_2d_mc(sh, &mapA, full_bar + s, kg * TC_BK, m0 + (int)crank * (TC_BM / 2), (uint16_t)3);
          }
        } else if (p.conv) {
          const int tap = kg / cpb, cb = kg % cpb;
          const int r = tap / 3, q = tap % 3;
          tma_load_4d(sa, &mapA, full_bar + s, cb * TC_BK, w0 + q - p.pad, h0 + r - p.pad, img);
        } else {
          tma_load_2d(sa, &mapA, full_bar + s, kg * TC_BK, m0);
        }
        if (dbg && kb == 0) p.dbg[2] = clock64();
        if (dbg && kb < 40) p.dbg[64 + kb] = clock64();          // A-producer: TMA of K block kb issued
      }
      if (dbg) p.dbg[3] = clock64();
    }
    __syncwarp();
  } else if (warp == 6) {
    if (lane == 0) {
      // ===== TMA producer, B tiles (weights) =====
      const uint64_t polB = p.w_evict_last ? l2_policy_evict_last() : l2_policy_evict_normal();
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        uint8_t* sb = smem + s * SM::STAGE_BYTES + SM::A_BYTES;
        mbar_expect_tx(full_bar + s, SM::B_BYTES);
        tma_load_2d_hint(sb, &mapB, full_bar + s, (kb0 + kb) * TC_BK, n0, polB);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (one thread) =====
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, NT);
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar + s, ph);
        if (dbg && kb == 0) p.dbg[4] = clock64();
        if (dbg && kb < 40) p.dbg[16 + kb] = clock64();          // MMA thread: K block kb landed
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint32_t sb = sa + SM::A_BYTES;
        const uint64_t da = make_kmajor_sw128_desc(sa);
        const uint64_t db = make_kmajor_sw128_desc(sb);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; k++) {
          // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the 16 B-unit address field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (MC) umma_commit_mc(empty_bar + s, (uint16_t)3);   // frees the slot in BOTH CTAs
        else umma_commit(empty_bar + s);                       // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);             // accumulator complete
      if (dbg) p.dbg[5] = clock64();
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5 ; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // row inside the 128-row tile == TMEM lane
    // while the main loop runs: stage the bias slice of this N tile in shared memory (a dependent global load per
    // 8 columns inside the drain loop cost ~6 us per tile before)
    {
      const int te = threadIdx.x - 64;
      if (te < NT) s_bias[te] = (p.bias && blockIdx.z == 0 && n0 + te < p.N) ? __ldg(p.bias + n0 + te) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    bool row_ok;
    int64_t row_off;
    if (p.conv) {
      const int h = h0 + row / p.BW, w = w0 + row % p.BW;
      row_ok = (h < p.Ho) && (w < p.Wo);
      row_off = (((int64_t)img * p.Ho + h) * p.Wo + w) * p.ldc;
    } else {
      row_ok = (m0 + row) < p.M;
      row_off = (int64_t)(m0 + row) * p.ldc;
    }
    const bool use_mask = p.mask && !p.out_f32 && !p.atomic;
    mbar_wait(tmem_full_bar, 0);
    if (dbg && warp == 2 && lane == 0) p.dbg[6] = clock64();
    tc_fence_after();
    if (p.lstm) {
      // ---- fused LSTM cell (nn.LSTMCell, gate order i,f,g,o).  The 32x32 accumulator chunk is transposed through the idle
      // pipeline ring so that a lane owns one hidden unit (4 interleaved gate columns) and consecutive lanes consecutive units
      float* stg = reinterpret_cast<float*>(smem) + quad * (32 * 36);
      const int D = p.l_D;
#pragma unroll 1
      for (int c = 0; c < NT; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int g = 0; g < 8; g++)
          *reinterpret_cast<uint4*>(&stg[lane * 36 + g * 4]) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        __syncwarp();
        const int u = lane & 7;
        const int j = (n0 + c) / 4 + u;
#pragma unroll 2
        for (int i = 0; i < 8; i++) {
          const int rl = i * 4 + (lane >> 3);
          const int b = m0 + quad * 32 + rl;
          if (b >= p.M || j >= D) continue;
          const float4 a = *reinterpret_cast<const float4*>(&stg[rl * 36 + 4 * u]);
          int64_t tk = p.l_tok[(int64_t)b * p.l_tok_stride];
          tk = tk < 0 ? 0 : (tk >= p.l_V ? p.l_V - 1 : tk);
          const float* pt = p.l_ptab + tk * 4 * D;
          const float* hh = p.l_hh + (int64_t)b * p.l_hh_stride;
          const float pi = a.x + pt[j] + hh[j];
          const float pf = a.y + pt[D + j] + hh[D + j];
          const float pg = a.z + pt[2 * D + j] + hh[2 * D + j];
          const float po = a.w + pt[3 * D + j] + hh[3 * D + j];
          const float ig = sigmoidf_(pi), fg = sigmoidf_(pf), gg = tanhf(pg), og = sigmoidf_(po);
          const float cn = fg * p.l_cprev[(int64_t)b * D + j] + ig * gg;
          const float hn = og * tanhf(cn);
          float* gt = p.l_gates + (int64_t)b * 4 * D;
          gt[j] = ig; gt[D + j] = fg; gt[2 * D + j] = gg; gt[3 * D + j] = og;
          p.l_c[(int64_t)b * D + j] = cn;
          p.l_h[(int64_t)b * D + j] = hn;
          if (p.l_hbf) p.l_hbf[(int64_t)b * D + j] = __float2bfloat16_rn(hn);
          if (p.l_hd) p.l_hd[(int64_t)b * p.l_hd_stride + j] = p.l_dmask ? hn * p.l_dmask[(int64_t)b * p.l_hd_stride + j] : hn;
        }
        __syncwarp();
      }
    } else
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
      uint4 mk4[4];
      if (use_mask && row_ok) {                 // issue the mask loads of the whole chunk before the TMEM read
#pragma unroll
        for (int g = 0; g < 4; g++)
          if (n0 + c + g * 8 < p.N) mk4[g] = *reinterpret_cast<const uint4*>(p.mask + row_off + n0 + c + g * 8);
      }
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);   // warp-collective: no divergence around it
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int n = n0 + c + g * 8;
        if (n >= p.N) continue;
        float f[8];
        const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[c + g * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[c + g * 8 + 4]);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = __uint_as_float(v[g * 8 + i]) + bb[i];
        if (p.atomic) {
          float* o = reinterpret_cast<float*>(p.out) + row_off + n;
#pragma unroll
          for (int i = 0; i < 8; i++) atomicAdd(o + i, f[i]);
        } else if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + row_off + n;
          if (p.accumulate) {
            float old[8];
            ld8(o, old);
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] += old[i];
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fmaxf(f[i], 0.f);
          }
          st8(o, f);
        } else {
          bf16* o = reinterpret_cast<bf16*>(p.out) + row_off + n;
          if (p.accumulate) {
            float old[8];
            ld8(o, old);
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] += old[i];
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fmaxf(f[i], 0.f);
          }
          if (use_mask) {
            const uint32_t w[4] = {mk4[g].x, mk4[g].y, mk4[g].z, mk4[g].w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              if (!(__uint_as_float(w[i] << 16) > 0.f)) f[2 * i] = 0.f;
              if (!(__uint_as_float(w[i] & 0xffff0000u) > 0.f)) f[2 * i + 1] = 0.f;
            }
          }
          st8(o, f);
        }
      }
    }
  }
  tc_fence_before();
  if (MC) cluster_sync_all(); else __syncthreads();   // a CTA may not exit while its peer can still multicast into it
  if (dbg && threadIdx.x == 0) p.dbg[7] = clock64();
  if (warp == 1) tmem_dealloc(tmem_base, NT);
  if (dbg && threadIdx.x == 32) p.dbg[8] = clock64();
}

// ------------------------------------------------------------------------------------------------
// Persistent 3x3 convolution (forward and data gradient): one CTA per SM walks over output tiles of 128 positions x NT
// channels; TWO TMEM accumulators (2 x NT columns) so that the 128-thread epilogue of tile i drains while the MMA thread
// already accumulates tile i+1; the TMA ring runs ahead across tile boundaries.  NT = 256 for the 256/512-channel layers:
// per 64-wide K block a CTA pulls 16 KB of A + 32 KB of B from L2 for 4.2 MFLOP = 85 FLOP per L2 byte (128 x 128 tiles: 64).
// The one-tile-per-CTA kernel above is L2-bandwidth bound at 45-56 % tensor-pipe activity (profiles/r1_tc_conv_r1_raw.csv).
// ------------------------------------------------------------------------------------------------
template <int NT, int STAGES, int MT>
struct TcPSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;              // one 128-position sub-tile
  static constexpr int B_BYTES = NT * TC_BK * 2;
  static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * NT * 4 /*bias, double buffered*/;
};

// MT = 2 (layers with <= 128 output channels): a CTA tile is TWO 128-position sub-tiles that share every weight (B) stage —
// FLOP per byte of A equals N in an implicit GEMM, so with N <= 128 the B stage is as large as an A tile and sharing it
// raises the FLOP per L2 byte from 64 to 87 (N = 128) / 43 to 52 (N = 64).
template <int NT, int STAGES, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_p_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB, TcParams p, int ntiles_n,
                                                                 int total_tiles, int mtiles) {
  using SM = TcPSmem<NT, STAGES, MT>;
  static_assert(2 * MT * NT <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;        // [2] accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;            // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * SM::STAGE_BYTES + 256);      // [2][NT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.K / TC_BK;
  const int cpb = p.Cin / TC_BK;
  const int per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar + s, 2);                // A producer + B producer
      mbar_init(empty_bar + s, 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(tfull_bar + b, 1);
      mbar_init(tempty_bar + b, 4);              // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * MT * NT);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer, A tiles (implicit im2col: tap = coordinate shift, halo = out-of-bounds zero fill) =====
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt0 = (tile / ntiles_n) * MT;
        const int nsub = min(MT, mtiles - mt0);
        int img[MT], h0[MT], w0[MT];
#pragma unroll
        for (int j = 0; j < MT; j++) {
          const int mt = min(mt0 + j, mtiles - 1);
          img[j] = mt / per_img;
          const int rem = mt % per_img;
          h0[j] = (rem / p.tiles_w) * p.BH;
          w0[j] = (rem % p.tiles_w) * p.BW;
        }
        for (int kb = 0; kb < KB; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          mbar_expect_tx(full_bar + s, (uint32_t)nsub * SM::A_BYTES);
          const int tap = kb / cpb, cb = kb % cpb;
          const int r = tap / 3, q = tap % 3;
#pragma unroll
          for (int j = 0; j < MT; j++)
            if (j < nsub)
              tma_load_4d(smem + s * SM::STAGE_BYTES + j * SM::A_BYTES, &mapA, full_bar + s, cb * TC_BK, w0[j] + q - p.pad,
                          h0[j] + r - p.pad, img[j]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 6) {
    if (lane == 0) {
      // ===== TMA producer, B tiles (weights: re-read by every CTA -> keep them in L2) =====
      const uint64_t polB = l2_policy_evict_last();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % ntiles_n) * NT;
        for (int kb = 0; kb < KB; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          mbar_expect_tx(full_bar + s, SM::B_BYTES);
          tma_load_2d_hint(smem + s * SM::STAGE_BYTES + MT * SM::A_BYTES, &mapB, full_bar + s, kb * TC_BK, n0, polB);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, NT);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, lt++) {
        const uint32_t buf = lt & 1;
        const int nsub = min(MT, mtiles - (tile / ntiles_n) * MT);
        mbar_wait(tempty_bar + buf, ((lt >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator set (2 tiles ago)
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * (MT * NT);
        for (int kb = 0; kb < KB; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          const uint64_t db = make_kmajor_sw128_desc(sa + MT * SM::A_BYTES);
#pragma unroll
          for (int j = 0; j < MT; j++) {
            if (j < nsub) {
              const uint64_t da = make_kmajor_sw128_desc(sa + j * SM::A_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; k++)
                umma_bf16(tacc + j * NT, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar + s);
        }
        umma_commit(tfull_bar + buf);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5 ; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int te = threadIdx.x - 64;                               // 0..127
    const bool use_mask = p.mask != nullptr;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, lt++) {
      const uint32_t buf = lt & 1;
      const int n0 = (tile % ntiles_n) * NT;
      const int mt0 = (tile / ntiles_n) * MT;
      const int nsub = min(MT, mtiles - mt0);
      float* sb = s_bias + buf * NT;
      for (int c = te; c < NT; c += 128) sb[c] = (p.bias && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");               // bias of this tile staged (buffer last read 2 tiles ago)
      mbar_wait(tfull_bar + buf, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < nsub; j++) {
        const int mt = mt0 + j;
        const int img = mt / per_img, rem = mt % per_img;
        const int h = (rem / p.tiles_w) * p.BH + row / p.BW, w = (rem % p.tiles_w) * p.BW + row % p.BW;
        const bool row_ok = (h < p.Ho) && (w < p.Wo);
        const int64_t row_off = (((int64_t)img * p.Ho + h) * p.Wo + w) * p.ldc;
        const uint32_t tacc = tmem_base + buf * (MT * NT) + j * NT + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < NT; c += 32) {
          uint4 mk4[4];
          if (use_mask && row_ok) {
#pragma unroll
            for (int g = 0; g < 4; g++)
              if (n0 + c + g * 8 < p.N) mk4[g] = *reinterpret_cast<const uint4*>(p.mask + row_off + n0 + c + g * 8);
          }
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)c, v);
          if (!row_ok) continue;
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int n = n0 + c + g * 8;
            if (n >= p.N) continue;
            float f[8];
            const float4 b0 = *reinterpret_cast<const float4*>(&sb[c + g * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sb[c + g * 8 + 4]);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = __uint_as_float(v[g * 8 + i]) + bb[i];
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; i++) f[i] = fmaxf(f[i], 0.f);
            }
            if (use_mask) {
              const uint32_t wv[4] = {mk4[g].x, mk4[g].y, mk4[g].z, mk4[g].w};
#pragma unroll
              for (int i = 0; i < 4; i++) {
                if (!(__uint_as_float(wv[i] << 16) > 0.f)) f[2 * i] = 0.f;
                if (!(__uint_as_float(wv[i] & 0xffff0000u) > 0.f)) f[2 * i + 1] = 0.f;
              }
            }
            st8(reinterpret_cast<bf16*>(p.out) + row_off + n, f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + buf);                  // this warp's quarter of the accumulators is free again
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * MT * NT);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:  dW[co][tap][ci] += sum_p dY[p][co] * X[p + tap][ci]
// GEMM view: M = co (128), N = ci (NT), K = output positions.  Both operands are "MN-major" (the channel index is
// contiguous in NHWC), which UMMA takes directly: smem tile = [128 positions][64 channels] (one TMA box, SWIZZLE_128B),
// descriptor LBO = distance between 64-channel boxes, SBO = 1024 B (8 positions), 16 positions (2048 B) per MMA.
// grid: (co tiles * ci tiles, 9 taps, K splits); epilogue = fp32 atomics into dW (split-K partial sums).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // next 64-element block along M/N
  d |= (uint64_t)(1024 >> 4) << 32;                   // next 8 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct WgParams {
  int Cin, Cout, Ho, Wo, pad;
  int BW, BH, tiles_w, tiles_h, kstages, per_split, ci_tiles;
  float* dw;
  int plain;          // 1: plain C[M][N] += A[K][M]^T B[K][N] (2-D maps; Cout = M, Cin = N, tap ignored)
                      // 2: the same, batched over blockIdx.y (3-D maps {M|N, K, batch}; C of batch b at dw + b * batch_c)
  int64_t ldc;
  int64_t batch_c;
};

template <int NT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS) tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapDY,
                                                               const __grid_constant__ CUtensorMap mapX, WgParams p) {
  constexpr int BOX = 128 * 64 * 2;                 // one [128 pos][64 ch] box
  constexpr int A_BYTES = 2 * BOX, B_BYTES = (NT / 64) * BOX, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = (blockIdx.x / p.ci_tiles) * 128, ci0 = (blockIdx.x % p.ci_tiles) * NT;
  const int tap = p.plain == 2 ? 0 : blockIdx.y, r = tap / 3, q = tap % 3;
  const int bz = p.plain == 2 ? blockIdx.y : 0;
  const int ks0 = blockIdx.z * p.per_split;
  const int ks1 = min(p.kstages, ks0 + p.per_split);
  const int KS = ks1 - ks0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapDY);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 2); mbar_init(empty_bar + s, 1); }   // two producers
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (KS > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int per_img = p.tiles_w * p.tiles_h;
        for (int i = 0; i < KS; i++) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          const int ks = ks0 + i;
          const int img = ks / per_img, rem = ks % per_img;
          const int h0 = (rem / p.tiles_w) * p.BH, w0 = (rem % p.tiles_w) * p.BW;
          uint8_t* sa = smem + s * STAGE_BYTES;
          mbar_expect_tx(full_bar + s, A_BYTES);
          if (p.plain == 2) {
            tma_load_3d(sa, &mapDY, full_bar + s, co0, ks * 128, bz);
            tma_load_3d(sa + BOX, &mapDY, full_bar + s, co0 + 64, ks * 128, bz);
          } else if (p.plain) {
            tma_load_2d(sa, &mapDY, full_bar + s, co0, ks * 128);
            tma_load_2d(sa + BOX, &mapDY, full_bar + s, co0 + 64, ks * 128);
          } else {
            tma_load_4d(sa, &mapDY, full_bar + s, co0, w0, h0, img);
            tma_load_4d(sa + BOX, &mapDY, full_bar + s, co0 + 64, w0, h0, img);
          }
        }
      }
      __syncwarp();
    } else if (warp == 6) {
      if (lane == 0) {            // second producer: the X (B operand) boxes
        const int per_img = p.tiles_w * p.tiles_h;
        for (int i = 0; i < KS; i++) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          const int ks = ks0 + i;
          const int img = ks / per_img, rem = ks % per_img;
          const int h0 = (rem / p.tiles_w) * p.BH, w0 = (rem % p.tiles_w) * p.BW;
          uint8_t* sb = smem + s * STAGE_BYTES + A_BYTES;
          mbar_expect_tx(full_bar + s, B_BYTES);
          if (p.plain == 2) {
#pragma unroll
            for (int j = 0; j < NT / 64; j++) tma_load_3d(sb + j * BOX, &mapX, full_bar + s, ci0 + 64 * j, ks * 128, bz);
          } else if (p.plain) {
#pragma unroll
            for (int j = 0; j < NT / 64; j++) tma_load_2d(sb + j * BOX, &mapX, full_bar + s, ci0 + 64 * j, ks * 128);
          } else {
#pragma unroll
            for (int j = 0; j < NT / 64; j++)
              tma_load_4d(sb + j * BOX, &mapX, full_bar + s, ci0 + 64 * j, w0 + q - p.pad, h0 + r - p.pad, img);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, NT) | (1u << 15) | (1u << 16);   // A and B MN-major
        for (int i = 0; i < KS; i++) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t da = make_mnmajor_sw128_desc(sa + k * 2048, BOX);
            const uint64_t db = make_mnmajor_sw128_desc(sb + k * 2048, BOX);
            umma_bf16(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar + s);
        }
        umma_commit(tmem_full_bar);
      }
      __syncwarp();
    } else {
      const int quad = warp & 3;
      float* stg = reinterpret_cast<float*>(smem) + quad * (32 * 36);    // transpose staging in the (now idle) ring
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < NT; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int g = 0; g < 8; g++)
          *reinterpret_cast<uint4*>(&stg[lane * 36 + g * 4]) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        __syncwarp();
        // one warp instruction = one 128 B row segment of dW -> coalesced red.global.add
#pragma unroll 4
        for (int rr = 0; rr < 32; rr++) {
          const int co = co0 + quad * 32 + rr;
          const int ci = ci0 + c + lane;
          if (co < p.Cout && ci < p.Cin) {
            float* o = p.plain ? p.dw + (int64_t)bz * p.batch_c + (int64_t)co * p.ldc + ci : p.dw + ((int64_t)co * 9 + tap) * p.Cin + ci;
            atomicAdd(o, stg[rr * 36 + lane]);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, NT);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_encode = nullptr;
static int g_tc_state = -1;   // -1 unknown, 0 unavailable, 1 ok

bool tc_available() {
  if (g_tc_state >= 0) return g_tc_state == 1;
  g_tc_state = 0;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) return false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
      qres != cudaDriverEntryPointSuccess)
    return false;
  g_encode = (PFN_tmapEncodeTiled)fn;
  g_tc_state = 1;
  return true;
}

// cuTensorMapEncodeTiled costs ~1-2 us of host time; the decoder issues ~600 skinny GEMMs per step with a handful
// of distinct (pointer, shape) combinations per step index, so encoded maps are cached.
struct MapKey {
  const void* base; int rank; cuuint64_t dims[4]; cuuint64_t str[3]; cuuint32_t box[4];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; i++) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
// the ABI allows one host thread per device: the descriptor cache is shared by all of them -> guarded
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static std::mutex g_map_mutex;

static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank;
  for (int i = 0; i < rank; i++) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; i++) key.str[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *m = it->second; return LO_OK; }
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LO_ECUDA, "%s: cuTensorMapEncodeTiled failed (%ld)", "tc", (long)r);
  std::lock_guard<std::mutex> lk(g_map_mutex);
  if (g_map_cache.size() > 20000) g_map_cache.clear();
  g_map_cache.emplace(key, *m);
  return LO_OK;
}

int g_opt_conv_mc = 1;    // cluster-of-2 multicast of the A tile
long long* g_tc_dbg = nullptr;
int g_opt_skinny8 = 1;    // 8-stage (198 KB smem) config for M <= 128 GEMMs; off when two decoder chains share the SMs

template <int NT, int STAGES, int MC>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, int mtiles, int splits, cudaStream_t st) {
  using SM = TcSmem<NT, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    LO_CUDA(cudaFuncSetAttribute(tc_gemm_conv_kernel<NT, STAGES, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  dim3 grid(cdiv(p.N, NT), mtiles, splits);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (g_opt_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  if (MC) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    n++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  LO_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_conv_kernel<NT, STAGES, MC>, mA, mB, p));
  LO_LAUNCH_OK();
  return LO_OK;
}

static int launch_tc_any(const CUtensorMap& mA, const CUtensorMap& mB, TcParams& p, int mtiles, int splits, int nt, cudaStream_t st,
                         int mc = 0) {
  const int KB = p.K / TC_BK;
  if (splits < 1) splits = 1;
  if (splits > KB) splits = KB;
  p.kb_per_split = cdiv(KB, splits);
  splits = cdiv(KB, p.kb_per_split);
  p.atomic = splits > 1 ? 1 : p.atomic;
  if (nt == 64 && mtiles == 1 && p.kb_per_split > 4 && g_opt_skinny8) return launch_tc<64, 8, 0>(mA, mB, p, mtiles, splits, st);   // skinny: all K in flight
  if (nt == 64) return launch_tc<64, 4, 0>(mA, mB, p, mtiles, splits, st);
  if (mc) return launch_tc<128, 3, 1>(mA, mB, p, mtiles, splits, st);
  return launch_tc<128, 3, 0>(mA, mB, p, mtiles, splits, st);
}

static TcLstmEpi g_lstm_epi;
static bool g_lstm_epi_on = false;

// splits > 1 (or atomic_acc): fp32 C only, partial sums are ADDED onto C with atomics (C must hold the base values)
int tc_gemm_nt_ex(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* C, int dtC, int64_t ldc, int M, int N, int K,
                  const float* bias, int accumulate, int relu, int splits, int atomic_acc, int small_n_tile, cudaStream_t st) {
  if (!tc_available()) return fail(LO_ENOTSUP, "%s: needs an sm_100 device", __func__);
  LO_CHECK_ARG(K % 64 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && ldc >= (N + 7) / 8 * 8, "K%64, ld%8, ldc >= roundup8(N)");
  LO_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0, "16-byte alignment");
  CUtensorMap mA, mB;
  const int NT = (N <= 64 || small_n_tile) ? 64 : 128;
  const int mc = (g_opt_conv_mc && NT == 128 && cdiv(N, 128) % 2 == 0 && splits <= 1 && !atomic_acc) ? 1 : 0;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t str[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(mc ? 64 : 128)};
    LO_TRY(make_map(&mA, A, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t str[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)NT};
    LO_TRY(make_map(&mB, W, 2, dims, str, box));
  }
  LO_CHECK_ARG(!(splits > 1 || atomic_acc) || (dtC == LO_F32 && !relu), "split-K needs fp32 output without ReLU");
  TcParams p{};
  p.M = M; p.N = N; p.K = K; p.conv = 0;
  p.bias = bias; p.mask = nullptr; p.out = C; p.ldc = ldc;
  p.out_f32 = (dtC == LO_F32); p.accumulate = accumulate; p.relu = relu; p.atomic = atomic_acc;
  p.w_evict_last = (M <= 128) ? 1 : 0;
  p.dbg = g_tc_dbg;
  if (g_lstm_epi_on) {
    const TcLstmEpi& e = g_lstm_epi;
    p.lstm = 1;
    p.l_ptab = e.ptab; p.l_tok = e.tok; p.l_tok_stride = e.tok_stride; p.l_hh = e.hh; p.l_hh_stride = e.hh_stride;
    p.l_cprev = e.c_prev; p.l_gates = e.gates; p.l_c = e.c_out; p.l_h = e.h_out; p.l_hbf = e.h_bf;
    p.l_hd = e.hd; p.l_hd_stride = e.hd_stride; p.l_dmask = e.dmask; p.l_D = e.D; p.l_V = e.V;
  }
  return launch_tc_any(mA, mB, p, cdiv(M, TC_BM), splits, NT, st, mc);
}

// gates = A @ Wil^T (Wil gate-interleaved [4D][K]) followed by the LSTM cell in the epilogue (see TcParams)
int tc_gemm_nt_lstm(const bf16* A, int64_t lda, const bf16* Wil, int64_t ldw, int M, int D, int K, const TcLstmEpi& e, cudaStream_t st) {
  g_lstm_epi = e;
  g_lstm_epi_on = true;
  const int r = tc_gemm_nt_ex(A, lda, Wil, ldw, e.gates /*unused as C*/, LO_F32, 4 * D, M, 4 * D, K, nullptr, 0, 0, 1, 0, 1, st);
  g_lstm_epi_on = false;
  return r;
}

int tc_gemm_nt(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* C, int dtC, int64_t ldc, int M, int N, int K,
               const float* bias, int accumulate, int relu, cudaStream_t st) {
  return tc_gemm_nt_ex(A, lda, W, ldw, C, dtC, ldc, M, N, K, bias, accumulate, relu, 1, 0, 0, st);
}

int g_opt_wgrad256 = 0;       // conv weight gradient with 128 x 256 tiles when Cin % 256 == 0: measured slower (two 96 KB stages
                              // cannot hide the loads: 3.64 vs 3.42 ms for the conv tensor kernels, run 50)
int g_opt_conv_persist = 1;   // persistent double-accumulator kernel (tc_conv_p_kernel); 0 = one tile per CTA (tc_gemm_conv_kernel)

int g_opt_conv_mt2 = 1;       // layers with <= 128 output channels: two 128-position sub-tiles per CTA tile share each weight stage

template <int NT, int STAGES, int MT>
static int launch_conv_p(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, int mtiles, cudaStream_t st) {
  using SM = TcPSmem<NT, STAGES, MT>;
  static bool attr_set = false;
  static int n_sm = 0;
  if (!attr_set) {
    LO_CUDA(cudaFuncSetAttribute(tc_conv_p_kernel<NT, STAGES, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    int dev = 0;
    LO_CUDA(cudaGetDevice(&dev));
    LO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int ntn = cdiv(p.N, NT);
  const int total = cdiv(mtiles, MT) * ntn;
  const int grid = total < n_sm ? total : n_sm;
  LO_CUDA(launch_pdl(tc_conv_p_kernel<NT, STAGES, MT>, dim3(grid), dim3(TC_THREADS), (size_t)SM::TOTAL, st, mA, mB, p, ntn, total, mtiles));
  LO_LAUNCH_OK();
  return LO_OK;
}

int tc_conv3x3(const bf16* x, const bf16* w, const float* bias, const bf16* mask, bf16* y, int N, int H, int W, int Cin, int Cout,
               int pad, int relu, cudaStream_t st) {
  if (!tc_available()) return fail(LO_ENOTSUP, "%s: needs an sm_100 device", __func__);
  LO_CHECK_ARG(Cin % 64 == 0 && Cout % 8 == 0, "Cin%64==0, Cout%8==0");
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  int BW = 128;
  while (BW > 8 && BW / 2 >= Wo) BW /= 2;     // smallest power of two >= Wo (capped at 128)
  const int BH = 128 / BW;
  if (g_opt_conv_persist) {
    const int NTp = Cout % 256 == 0 ? 256 : (Cout > 64 ? 128 : 64);
    CUtensorMap mA, mB;
    {
      cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
      LO_TRY(make_map(&mA, x, 4, dims, str, box));
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
      cuuint64_t str[1] = {(cuuint64_t)9 * Cin * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)NTp};
      LO_TRY(make_map(&mB, w, 2, dims, str, box));
    }
    TcParams p{};
    p.M = N * Ho * Wo; p.N = Cout; p.K = 9 * Cin; p.conv = 1;
    p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.pad = pad;
    p.BW = BW; p.BH = BH; p.tiles_w = cdiv(Wo, BW); p.tiles_h = cdiv(Ho, BH);
    p.bias = bias; p.mask = mask; p.out = y; p.ldc = Cout; p.relu = relu;
    const int mtiles = N * p.tiles_w * p.tiles_h;
    if (NTp == 256) return launch_conv_p<256, 4, 1>(mA, mB, p, mtiles, st);
    if (NTp == 128) return g_opt_conv_mt2 ? launch_conv_p<128, 4, 2>(mA, mB, p, mtiles, st) : launch_conv_p<128, 6, 1>(mA, mB, p, mtiles, st);
    return g_opt_conv_mt2 ? launch_conv_p<64, 4, 2>(mA, mB, p, mtiles, st) : launch_conv_p<64, 8, 1>(mA, mB, p, mtiles, st);
  }
  CUtensorMap mA, mB;
  const int NT = Cout <= 64 ? 64 : 128;
  const int mc = (g_opt_conv_mc && NT == 128 && cdiv(Cout, 128) % 2 == 0) ? 1 : 0;
  // MC: each CTA of the pair loads one 64-position half of the tile (the first BH/2 rows, or the first BW/2 columns when BH == 1)
  const int hbw = mc ? (BH >= 2 ? BW : BW / 2) : BW, hbh = mc ? (BH >= 2 ? BH / 2 : 1) : BH;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)hbw, (cuuint32_t)hbh, 1};
    LO_TRY(make_map(&mA, x, 4, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t str[1] = {(cuuint64_t)9 * Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)NT};
    LO_TRY(make_map(&mB, w, 2, dims, str, box));
  }
  TcParams p{};
  p.M = N * Ho * Wo; p.N = Cout; p.K = 9 * Cin; p.conv = 1;
  p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.pad = pad;
  p.BW = BW; p.BH = BH; p.tiles_w = cdiv(Wo, BW); p.tiles_h = cdiv(Ho, BH);
  p.bias = bias; p.mask = mask; p.out = y; p.ldc = Cout;
  p.out_f32 = 0; p.accumulate = 0; p.relu = relu;
  p.dbg = g_tc_dbg;
  p.half_w = (mc && BH < 2) ? BW / 2 : 0;
  p.half_h = (mc && BH >= 2) ? BH / 2 : 0;
  const int mtiles = N * p.tiles_w * p.tiles_h;
  LO_CHECK_ARG(mtiles <= 65535, "too many M tiles for grid.y");
  return launch_tc_any(mA, mB, p, mtiles, 1, NT, st, mc);
}


template <int NT, int STAGES>
static int launch_wgrad(const CUtensorMap& mDY, const CUtensorMap& mX, const WgParams& p, dim3 grid, cudaStream_t st) {
  constexpr int TOTAL = STAGES * (2 + NT / 64) * 16384 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    LO_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
    attr_set = true;
  }
  tc_wgrad_kernel<NT, STAGES><<<grid, TC_THREADS, TOTAL, st>>>(mDY, mX, p);
  LO_LAUNCH_OK();
  return LO_OK;
}

// dw must be zeroed by the caller (split-K partial sums are accumulated with atomics)
int tc_conv3x3_wgrad(const bf16* x, const bf16* dy, float* dw, int N, int H, int W, int Cin, int Cout, int pad, cudaStream_t st) {
  if (!tc_available()) return fail(LO_ENOTSUP, "%s: needs an sm_100 device", __func__);
  LO_CHECK_ARG(Cin % 64 == 0 && Cout % 128 == 0, "Cin%64==0, Cout%128==0");
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  int BW = 128;
  while (BW > 8 && BW / 2 >= Wo) BW /= 2;
  const int BH = 128 / BW;
  CUtensorMap mDY, mX;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wo * Cout * 2, (cuuint64_t)Ho * Wo * Cout * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    LO_TRY(make_map(&mDY, dy, 4, dims, str, box));
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    LO_TRY(make_map(&mX, x, 4, dims, str, box));
  }
  WgParams p{};
  p.Cin = Cin; p.Cout = Cout; p.Ho = Ho; p.Wo = Wo; p.pad = pad;
  p.BW = BW; p.BH = BH; p.tiles_w = cdiv(Wo, BW); p.tiles_h = cdiv(Ho, BH);
  p.kstages = N * p.tiles_w * p.tiles_h;
  p.dw = dw;
  // 128 (co) x 256 (ci) tiles when Cin allows it: 87 instead of 64 FLOP per byte pulled from L2 (two 96 KB stages)
  const int NT = (g_opt_wgrad256 && Cin % 256 == 0) ? 256 : (Cin >= 128 ? 128 : 64);
  p.ci_tiles = Cin / NT;
  const int tiles = (Cout / 128) * p.ci_tiles * 9;
  int splits = cdiv(NT == 256 ? 148 : 148 * 2, tiles);
  if (splits > p.kstages) splits = p.kstages;
  if (splits < 1) splits = 1;
  p.per_split = cdiv(p.kstages, splits);
  splits = cdiv(p.kstages, p.per_split);
  dim3 grid((Cout / 128) * p.ci_tiles, 9, splits);
  if (NT == 256) return launch_wgrad<256, 2>(mDY, mX, p, grid, st);
  if (NT == 128) return launch_wgrad<128, 3>(mDY, mX, p, grid, st);
  return launch_wgrad<64, 4>(mDY, mX, p, grid, st);
}

// C[M][N] (fp32, ldc) += A[K][M]^T * B[K][N]; A, B bf16 row-major.  C must hold the values to accumulate onto
// (zero it for a plain product): split-K partial sums land with fp32 atomics.
int tc_gemm_tn(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  if (!tc_available()) return fail(LO_ENOTSUP, "%s: needs an sm_100 device", __func__);
  LO_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "lda%8, ldb%8");
  LO_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "16-byte alignment");
  CUtensorMap mA, mB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)K};
    cuuint64_t str[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {64, 128};
    LO_TRY(make_map(&mA, A, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)K};
    cuuint64_t str[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t box[2] = {64, 128};
    LO_TRY(make_map(&mB, B, 2, dims, str, box));
  }
  WgParams p{};
  p.plain = 1; p.Cout = M; p.Cin = N; p.dw = C; p.ldc = ldc;
  p.tiles_w = p.tiles_h = 1; p.BW = 128; p.BH = 1;
  p.kstages = cdiv(K, 128);
  const int NT = N > 64 ? 128 : 64;
  p.ci_tiles = cdiv(N, NT);
  const int mt = cdiv(M, 128);
  const int tiles = mt * p.ci_tiles;
  int splits = cdiv(148 * 2, tiles);
  if (splits > p.kstages) splits = p.kstages;
  if (splits < 1) splits = 1;
  p.per_split = cdiv(p.kstages, splits);
  splits = cdiv(p.kstages, p.per_split);
  dim3 grid(mt * p.ci_tiles, 1, splits);
  if (NT == 128) return launch_wgrad<128, 3>(mA, mB, p, grid, st);
  return launch_wgrad<64, 4>(mA, mB, p, grid, st);
}

// Batched C[b][M][N] (fp32, += with atomics) += A[b][K][M]^T * B[b][K][N]; A, B bf16 with arbitrary (16-byte multiple) strides:
// element (b, k, m) of A at A + b * sAb + k * sAk + m, likewise B.  One launch, grid.y = batch.  (decoder backward:
// d enc[b] += alphas[b]^T dctx[:, b, :], the context read summed over time.)
int tc_gemm_tn_batched(const bf16* A, int64_t sAk, int64_t sAb, const bf16* B, int64_t sBk, int64_t sBb, float* C, int64_t ldc,
                       int64_t sCb, int M, int N, int K, int batch, cudaStream_t st) {
  if (!tc_available()) return fail(LO_ENOTSUP, "%s: needs an sm_100 device", __func__);
  LO_CHECK_ARG(sAk % 8 == 0 && sAb % 8 == 0 && sBk % 8 == 0 && sBb % 8 == 0, "strides must be multiples of 8 elements");
  LO_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && batch >= 1 && batch <= 65535, "alignment / batch");
  CUtensorMap mA, mB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)K, (cuuint64_t)batch};
    cuuint64_t str[2] = {(cuuint64_t)sAk * 2, (cuuint64_t)sAb * 2};
    cuuint32_t box[3] = {64, 128, 1};
    LO_TRY(make_map(&mA, A, 3, dims, str, box));
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)K, (cuuint64_t)batch};
    cuuint64_t str[2] = {(cuuint64_t)sBk * 2, (cuuint64_t)sBb * 2};
    cuuint32_t box[3] = {64, 128, 1};
    LO_TRY(make_map(&mB, B, 3, dims, str, box));
  }
  WgParams p{};
  p.plain = 2; p.Cout = M; p.Cin = N; p.dw = C; p.ldc = ldc; p.batch_c = sCb;
  p.tiles_w = p.tiles_h = 1; p.BW = 128; p.BH = 1;
  p.kstages = cdiv(K, 128);
  p.per_split = p.kstages;                      // no split-K: the batch dimension supplies the parallelism
  const int NT = N > 64 ? 128 : 64;
  p.ci_tiles = cdiv(N, NT);
  dim3 grid(cdiv(M, 128) * p.ci_tiles, batch, 1);
  if (NT == 128) return launch_wgrad<128, 3>(mA, mB, p, grid, st);
  return launch_wgrad<64, 4>(mA, mB, p, grid, st);
}

}  // namespace lo
