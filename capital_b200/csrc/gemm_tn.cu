// FP64 tensor-core GEMM for sm_100a:  C[m x n] = alpha * A^T B + beta * C   (A: k x m, B: k x n, col-major)
//
// This is the kernel behind every trailing update of the CholInv schedule -- the reference's
// cblas_dgemm(T,N) in summa::syrk_internal (summa.hpp:143-145), cblas_dtrmm in summa::invoke
// (summa.hpp:64,71) and the Gram products of cacqr (cacqr.hpp:15,95) -- with the triangular
// structure the reference throws away (summa.hpp:115-116) turned into skipped k-tiles / output tiles.
//
// B200 design.  FP64 has no tcgen05 path (ptxas: "Unknown modifier .kind::f64"); the FP64 tensor pipe
// is reached through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; measured 37.2 TFLOP/s = 64 FMA/clk/SM, see
// profiles/r01_fp64_pipe_ceilings.log).  Both operands are K-contiguous, so a (rows x 16 k) tile is one
// 128-byte row per matrix column: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages it with one
// instruction per operand per stage from a dedicated producer warp; consumers wait on mbarriers (no
// __syncthreads in the main loop).  Inside a 16-wide k tile the four DMMAs use the k permutation
// {4q+j}: lane q then owns 32 contiguous bytes of every row, read as two conflict-free LDS.128
// (the 128B swizzle XORs the 16B chunk index with row%8, so the 8 lanes of a quarter-warp -- two rows x
// four q -- hit 8 distinct chunks).  Out-of-range rows/columns are zero-filled by TMA, so ragged M/N/K
// need no predicates in the main loop.
#include "common.cuh"

namespace {

struct GemmParams {
  int M, N, K;
  int rowoffA, rowoffB;  // element offset of the operand's first row inside its (16B-aligned) tensor map
  int flags;
  int noff;    // column index of this window of B / C inside the full operand (column-chunked launches)
  int koff;    // row index of this window inside the full operand (k-chunked launches keep the triangular k ranges right)
  int ksplit;  // gridDim.z chunks of the k range; > 1 => epilogue accumulates with atomics (C pre-initialised, beta ignored)
  double alpha, beta;
  double* C;
  long long ldc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int BK = 16;  // doubles per k tile = one 128-byte swizzle row

// Warp roles: NCW consumer warps (whole warpgroups) + one producer warpgroup of which a single lane drives TMA.
// Registers are allocated per warpgroup on sm_100, so a 9th warp would be charged as four anyway; with RC > 0 the
// producer group hands its registers to the consumers (setmaxnreg), which is what lets a 64x32 warp tile
// (128 accumulator registers) live without spills.
template <int BM, int BN, int WM, int WN, int STAGES, int MINB, int RC, int RP>
__global__ void __launch_bounds__(((BM / WM) * (BN / WN) + 4) * 32, MINB)
    gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmParams p) {
  constexpr int NWM = BM / WM, NWN = BN / WN, NCW = NWM * NWN;
  constexpr int FM = WM / 8, FN = WN / 8;
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;

  const int flags = p.flags;
  // Longest-tile-first over the WHOLE grid: with a triangular operand the k extent depends on one tile coordinate only, so the
  // linear CTA id is mapped to tiles in order of decreasing k extent (row-major over the other coordinate).  The first wave then
  // holds the longest tiles and second residents / the tail get the short ones (a per-column reversal alone interleaves long and
  // short tiles and lets two long tiles share an SM).
  const int lin = (int)(blockIdx.x + gridDim.x * blockIdx.y);
  int tm, tn;
  if (flags & CAPITAL_GEMM_A_UPPER) { tm = (int)gridDim.x - 1 - lin / (int)gridDim.y; tn = lin % (int)gridDim.y; }
  else if (flags & CAPITAL_GEMM_B_UPPER) { tn = (int)gridDim.y - 1 - lin / (int)gridDim.x; tm = lin % (int)gridDim.x; }
  else if (flags & CAPITAL_GEMM_A_LOWER) { tm = lin / (int)gridDim.y; tn = lin % (int)gridDim.y; }
  else if (flags & CAPITAL_GEMM_B_LOWER) { tn = lin / (int)gridDim.x; tm = lin % (int)gridDim.x; }
  else { tm = (int)blockIdx.x; tn = (int)blockIdx.y; }
  const int m0 = tm * BM, n0 = tn * BN;
  const int n0g = n0 + p.noff;  // column position used by the structure tests
  if ((flags & CAPITAL_GEMM_C_UPPER) && m0 > n0g + BN - 1) return;  // tile strictly below the diagonal

  int kb = 0, ke = p.K;
  if (flags & CAPITAL_GEMM_A_UPPER) ke = min(ke, m0 + BM - p.koff);
  if (flags & CAPITAL_GEMM_A_LOWER) kb = max(kb, m0 - p.koff);
  if (flags & CAPITAL_GEMM_B_UPPER) ke = min(ke, n0g + BN - p.koff);
  if (flags & CAPITAL_GEMM_B_LOWER) kb = max(kb, n0g - p.koff);
  kb &= ~(BK - 1);
  int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;
  if (nk == 0 && p.beta == 1.0 && p.ksplit <= 1) return;  // nothing to add (k-chunk entirely outside the operand's triangle)
  if (p.ksplit > 1) {  // this CTA's contiguous chunk of k tiles
    const int per = (nk + p.ksplit - 1) / p.ksplit;
    const int t0 = min(nk, (int)blockIdx.z * per), t1 = min(nk, t0 + per);
    kb += t0 * BK;
    nk = t1 - t0;
    if (nk == 0) return;
  }

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle, computed as an OFFSET into the shared array so that the pointer keeps its
  // shared address space (fragment loads then compile to LDS.128 instead of generic LD.E.128)
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_base + STAGES * STAGE_BYTES;
  const uint32_t empty0 = full0 + STAGES * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // ---------------- TMA producer warpgroup ----------------
    if (RC > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RP));
    if (warp == NCW && lane == 0) {
      for (int it = 0; it < nk; it++) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty0 + s * 8, ph ^ 1);
        mbar_expect_tx(full0 + s * 8, STAGE_BYTES);
        const int kk = kb + it * BK;
        tma_load_2d(smem_base + s * STAGE_BYTES, &mapA, full0 + s * 8, p.rowoffA + kk, m0);
        tma_load_2d(smem_base + s * STAGE_BYTES + A_BYTES, &mapB, full0 + s * 8, p.rowoffB + kk, n0);
      }
    }
    return;
  }

  // ---------------- DMMA consumers ----------------
  if (RC > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RC));
  const int wm = warp % NWM, wn = warp / NWM;
  const int g = lane >> 2, q = lane & 3;
  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int a_row_off = (wm * WM + g) * 128;
  const int b_row_off = A_BYTES + (wn * WN + g) * 128;
  const int swz[2] = {((2 * q) ^ g) * 16, ((2 * q + 1) ^ g) * 16};

  for (int it = 0; it < nk; it++) {
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1;
    mbar_wait(full0 + s * 8, ph);
    const uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double2 af[FM], bf[FN];
#pragma unroll
      for (int i = 0; i < FM; i++) af[i] = *reinterpret_cast<const double2*>(st + a_row_off + i * 1024 + swz[h]);
#pragma unroll
      for (int j = 0; j < FN; j++) bf[j] = *reinterpret_cast<const double2*>(st + b_row_off + j * 1024 + swz[h]);
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + s * 8);
  }

  // ---------------- epilogue: C = alpha * acc + beta * C ----------------
  const double alpha = p.alpha, beta = p.beta;
  const bool upper_only = flags & CAPITAL_GEMM_C_UPPER;
#pragma unroll
  for (int j = 0; j < FN; j++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int col = n0 + wn * WN + j * 8 + 2 * q + e;
      if (col >= p.N) continue;
      double* cc = p.C + (long long)col * p.ldc;
#pragma unroll
      for (int i = 0; i < FM; i++) {
        const int row = m0 + wm * WM + i * 8 + g;
        if (row >= p.M || (upper_only && row > col + p.noff)) continue;
        double v = alpha * acc[i][j][e];
        if (p.ksplit > 1) { atomicAdd(cc + row, v); continue; }
        if (beta != 0.0) v += beta * cc[row];
        cc[row] = v;
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, int STAGES, int MINB, int RC, int RP>
struct GemmCfg {
  static_assert(((BM / WM) * (BN / WN)) % 4 == 0, "consumer warps must form whole warpgroups");
  static constexpr int threads = ((BM / WM) * (BN / WN) + 4) * 32;
  static constexpr int smem = STAGES * (BM + BN) * 128 + 2 * STAGES * 8 + 1024;
  static constexpr auto kernel = gemm_tn_kernel<BM, BN, WM, WN, STAGES, MINB, RC, RP>;
};
using CfgBig = GemmCfg<128, 128, 64, 32, 5, 1, 232, 40>;   // 8 consumer warps + producer group, 1 CTA / SM
using CfgSmall = GemmCfg<64, 64, 32, 32, 6, 2, 0, 0>;       // 4 consumer warps + producer group, 2 CTAs / SM

capital_status_t make_map(capital_ctx* ctx, CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld,
                          int box_rows_k, int box_cols) {
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)box_rows_k, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled failed: CUresult " + std::to_string((int)r) + " rows=" + std::to_string(rows) +
                   " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld));
    return CAPITAL_ERR_CUDA;
  }
  return CAPITAL_OK;
}

template <class Cfg, int BM, int BN>
capital_status_t launch(capital_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                        int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int ksplit, int koff = 0, int noff = 0) {
  static bool attr_set = false;
  if (!attr_set) {
    CAP_CUDA(cudaFuncSetAttribute(Cfg::kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem));
    attr_set = true;
  }
  GemmParams p;
  p.M = (int)m; p.N = (int)n; p.K = (int)k; p.flags = flags; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.ksplit = ksplit; p.koff = koff; p.noff = noff;
  // TMA fetches 16-byte granules: a window that starts on an odd row (8-byte aligned only) cannot be addressed by
  // box coordinates, so it is first copied to an aligned scratch (O(k m) bytes against O(k m n) flops; only odd
  // split points of non-power-of-two sizes ever take this path).
  const bool same = (A == B && lda == ldb && m == n);
  if ((uintptr_t)A & 15) {
    double* sc;
    const int64_t lds = round_up(k, 2);
    CAP_TRY(ctx->workspace(st == ctx->side ? "gemm_alignA_side" : "gemm_alignA", (size_t)lds * m * 8, (void**)&sc));
    CAP_TRY(copy_block(ctx, st, k, m, A, lda, sc, lds));
    if (same) { B = sc; ldb = lds; }
    A = sc; lda = lds;
  }
  if ((uintptr_t)B & 15) {
    double* sc;
    const int64_t lds = round_up(k, 2);
    CAP_TRY(ctx->workspace(st == ctx->side ? "gemm_alignB_side" : "gemm_alignB", (size_t)lds * n * 8, (void**)&sc));
    CAP_TRY(copy_block(ctx, st, k, n, B, ldb, sc, lds));
    B = sc; ldb = lds;
  }
  p.rowoffA = 0;
  p.rowoffB = 0;
  CUtensorMap mapA, mapB;
  CAP_TRY(make_map(ctx, &mapA, A, k, m, lda, BK, BM));
  CAP_TRY(make_map(ctx, &mapB, B, k, n, ldb, BK, BN));
  dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(n, BN), (unsigned)ksplit);
  Cfg::kernel<<<grid, Cfg::threads, Cfg::smem, st>>>(mapA, mapB, p);
  CAP_CUDA(cudaGetLastError());
  return CAPITAL_OK;
}

}  // namespace

// Split-K variant for short-and-fat products (the tall-skinny Gram matrix, cacqr.hpp:15): C += alpha A^T B with the
// k range cut into `ksplit` chunks, partial tiles accumulated with FP64 atomics.  C must hold the addend on entry.
capital_status_t gemm_tn_splitk(capital_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                                int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int flags) {
  if (m <= 0 || n <= 0 || k <= 0) return CAPITAL_OK;
  if (lda < k || ldb < k || ldc < m || (lda & 1) || (ldb & 1)) return CAPITAL_ERR_INVALID;
  const int64_t tiles = ceil_div(m, 64) * ceil_div(n, 64);
  int64_t ks = ceil_div((int64_t)ctx->num_sms * 2, tiles);
  const int64_t max_ks = ceil_div(k, 16 * 32);  // at least 32 k-tiles per chunk
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  ctx->counters.kernel_launches++;
  ctx->counters.gemm_launches++;
  ctx->counters.gemm_flops += 2.0 * (double)m * (double)n * (double)k * ((flags & CAPITAL_GEMM_C_UPPER) ? 0.5 : 1.0);
  if (ks == 1) return launch<CfgSmall, 64, 64>(ctx, st, m, n, k, alpha, A, lda, B, ldb, 1.0, C, ldc, flags, 1);
  return launch<CfgSmall, 64, 64>(ctx, st, m, n, k, alpha, A, lda, B, ldb, 1.0, C, ldc, flags, (int)ks);
}

// Same product issued as a sequence of k-chunked launches (C accumulates).  Used for deferred work on the low-priority
// stream: a 128x128 tile with k = 8192 occupies its SM for ~1 ms, which would make the latency-critical kernels of the
// high-priority stream wait that long for an SM; chunks of `kc` bound the wait to kc/16 k-tiles.
capital_status_t gemm_tn_chunked(capital_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int64_t kc) {
  if (kc <= 0 || k <= kc + kc / 2) return gemm_tn(ctx, st, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags);
  for (int64_t k0 = 0; k0 < k; k0 += kc) {
    const int64_t kk = (k - k0 < kc + kc / 2) ? k - k0 : kc;
    CAP_TRY(gemm_tn_off(ctx, st, m, n, kk, alpha, A + k0, lda, B + k0, ldb, k0 == 0 ? beta : 1.0, C, ldc, flags, (int)k0, 0));
    if (kk != kc) break;
  }
  return CAPITAL_OK;
}

capital_status_t gemm_tn(capital_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                         int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags) {
  return gemm_tn_off(ctx, st, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags, 0, 0);
}

capital_status_t gemm_tn_off(capital_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                             int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, int koff, int noff) {
  if (m <= 0 || n <= 0) return CAPITAL_OK;
  if (k < 0 || lda < k || ldb < k || ldc < m || (lda & 1) || (ldb & 1) || ((uintptr_t)A & 7) || ((uintptr_t)B & 7)) {
    ctx->set_error("gemm_tn: invalid/unsupported leading dimensions (lda, ldb must be even and >= k)");
    return CAPITAL_ERR_INVALID;
  }
  if (k == 0) { ctx->set_error("gemm_tn: k must be positive"); return CAPITAL_ERR_INVALID; }
  if (m >= (1LL << 31) || n >= (1LL << 31) || k >= (1LL << 31) - 16) return CAPITAL_ERR_INVALID;
  ctx->counters.kernel_launches++;
  ctx->counters.gemm_launches++;
  // algorithmic flops of this product (structure exploited exactly, not tile-rounded)
  double f = 2.0 * (double)m * (double)n * (double)k;
  const bool atri = flags & (CAPITAL_GEMM_A_UPPER | CAPITAL_GEMM_A_LOWER), btri = flags & (CAPITAL_GEMM_B_UPPER | CAPITAL_GEMM_B_LOWER);
  if (atri && btri) f = 2.0 * (double)m * (double)n * (double)k / 3.0;
  else if (atri) f = (double)n * (double)m * (double)(m + 1);
  else if (btri && (flags & CAPITAL_GEMM_B_UPPER) && noff > 0 && koff == 0 && k >= noff + n) f = (double)m * (double)n * (double)(2 * (int64_t)noff + n + 1);
  else if (btri) f = (double)m * (double)n * (double)(n + 1);
  else if (flags & CAPITAL_GEMM_C_UPPER) f = (double)k * (double)m * (double)(m + 1);
  ctx->counters.gemm_flops += f;
  const int64_t tiles_big = ceil_div(m, 128) * ceil_div(n, 128);
  if (tiles_big >= ctx->num_sms) {
    // dominant kernel: optionally bracketed by events on its own stream (capital_profile_begin/end)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profiling) {
      CAP_TRY(ctx->prof_event(&e0)); CAP_TRY(ctx->prof_event(&e1));
      CAP_CUDA(cudaEventRecord(e0, st));
    }
    CAP_TRY((launch<CfgBig, 128, 128>(ctx, st, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags, 1, koff, noff)));
    if (ctx->profiling) {
      CAP_CUDA(cudaEventRecord(e1, st));
      ctx->prof_recs.push_back({e0, e1, f});
    }
    return CAPITAL_OK;
  }
  return launch<CfgSmall, 64, 64>(ctx, st, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags, 1, koff, noff);
}
