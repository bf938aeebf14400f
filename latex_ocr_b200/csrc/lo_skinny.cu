// Latency-optimised GEMM for the decoder's per-step projections: C[M <= 64][N] (+)= A[M][K] * W[N][K]^T + bias.
//
// These GEMMs are 33-134 MFLOP with M = batch = 64 rows: nothing about them is throughput bound.  The tcgen05 path
// (lo_tc.cu, kept and selectable with lo_set_option("skinny_mma", 0)) pays ~5.6 us per launch in fixed pipeline latency
// (TMEM allocation, single-thread TMA issue at ~430 cycles per K block, accumulator drain through 128 threads).  Here
// every CTA (16 output columns x all 64 rows x one K slice) issues ALL of its loads at once with cp.async (one L2
// round trip), then runs warp-level mma.sync from shared memory and stores straight from the accumulator registers.
// The FLOP-heavy work (convolutions, att1, weight gradients, logits) stays on tcgen05.
#include "lo_common.cuh"
#include "lo_ptx.cuh"

namespace lo {

int g_opt_skinny_mma = 1;

constexpr int SK_KC = 512;        // K elements per CTA
constexpr int SK_NT = 16;         // output columns per CTA
constexpr int SK_PITCH = SK_KC + 8;   // bf16 elements per smem row (+16 B: conflict-free ldmatrix)
constexpr int SK_SMEM = (64 + SK_NT) * SK_PITCH * 2;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;           // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// grid (N/16, ksplit), 128 threads (warp w owns rows 16w..16w+15)
__global__ void __launch_bounds__(128) skinny_mma_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W, int64_t ldw,
                                                          float* __restrict__ C, int64_t ldc, int M, int N, int K, int kc,
                                                          const float* __restrict__ bias, int atomic) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  bf16* sA = reinterpret_cast<bf16*>(sk_smem);               // [64][SK_PITCH]
  bf16* sW = sA + 64 * SK_PITCH;                             // [16][SK_PITCH]
  const int n0 = blockIdx.x * SK_NT;
  const int k0 = blockIdx.y * kc;
  const int kn = min(kc, K - k0);                            // multiple of 16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cpr = kn / 8;                                    // 16-byte chunks per row
  // PDL: the weight slice never depends on the preceding launch -> fetch it before griddepcontrol.wait
  for (int i = tid; i < SK_NT * cpr; i += 128) {
    const int r = i / cpr, c = i % cpr;
    cp_async16(sW + r * SK_PITCH + c * 8, W + (int64_t)min(n0 + r, N - 1) * ldw + k0 + c * 8, n0 + r < N);
  }
  pdl_wait();
  pdl_trigger();
  for (int i = tid; i < 64 * cpr; i += 128) {
    const int r = i / cpr, c = i % cpr;
    cp_async16(sA + r * SK_PITCH + c * 8, A + (int64_t)min(r, M - 1) * lda + k0 + c * 8, r < M);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const bf16* a_ptr = sA + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SK_PITCH + (lane >> 4) * 8;
  const bf16* b_ptr = sW + ((lane & 7) + (lane >> 4) * 8) * SK_PITCH + ((lane >> 3) & 1) * 8;
#pragma unroll 4
  for (int k = 0; k < kn; k += 16) {
    uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
    ldmatrix_x4(a0, a1, a2, a3, a_ptr + k);
    ldmatrix_x4(b0, b1, b2, b3, b_ptr + k);
    mma_bf16_16816(acc[0], a0, a1, a2, a3, b0, b1);          // columns n0 .. n0+7
    mma_bf16_16816(acc[1], a0, a1, a2, a3, b2, b3);          // columns n0+8 .. n0+15
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int n = n0 + j * 8 + 2 * t;
    if (n >= N) continue;
    const float bx = (bias && blockIdx.y == 0) ? bias[n] : 0.f, by = (bias && blockIdx.y == 0) ? bias[n + 1] : 0.f;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = warp * 16 + g + h * 8;
      if (r >= M) continue;
      float* o = C + (int64_t)r * ldc + n;
      const float x = acc[j][2 * h] + bx, y = acc[j][2 * h + 1] + by;
      if (atomic) { atomicAdd(o, x); atomicAdd(o + 1, y); }
      else *reinterpret_cast<float2*>(o) = make_float2(x, y);
    }
  }
}

// splits > 1 or atomic_acc: partial sums are added onto C with fp32 atomics (C holds the base values)
int skinny_gemm_nt(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                   int splits, int atomic_acc, cudaStream_t st) {
  LO_CHECK_ARG(M >= 1 && M <= 64 && K % 16 == 0 && N % 2 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 2 == 0, "M<=64, K%16, ld%8");
  static bool attr = false;
  if (!attr) {
    LO_CUDA(cudaFuncSetAttribute(skinny_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    attr = true;
  }
  int ks = cdiv(K, SK_KC);
  if (splits > ks) ks = splits;
  int kc = cdiv(cdiv(K, ks), 16) * 16;
  if (kc > SK_KC) kc = SK_KC;
  ks = cdiv(K, kc);
  const int atomic = (ks > 1 || atomic_acc) ? 1 : 0;
  if (ks > 1 && !atomic_acc) LO_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  LO_CUDA(launch_pdl(skinny_mma_kernel, dim3(cdiv(N, SK_NT), ks), dim3(128), (size_t)SK_SMEM, st, A, lda, W, ldw, C, ldc, M, N, K, kc, bias,
                     atomic));
  LO_LAUNCH_OK();
  return LO_OK;
}

}  // namespace lo
