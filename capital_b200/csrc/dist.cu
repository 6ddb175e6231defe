// Multi-GPU schedules: one process per GPU, NCCL over NVLink/NVSwitch in place of the reference's MPI.
//
//  * CholInv on the reference's c x d x d grid (c == d): every SUMMA of cholinv::invoke (summa.hpp:46-161) becomes
//      fetch X block from (y,z,z) and Y block from (x,z,z)   [util::transpose + row/column MPI_Bcast, summa.hpp:185,193]
//      local DMMA product on the k = z (mod d) slice          [cblas_dgemm / dtrmm,                   summa.hpp:64,143]
//      depth all-reduce of the partial result                 [MPI_Allreduce over depth,              summa.hpp:236]
//    and the base case is the replicate-everything policy (cholinv/policy.h:160-224): all-gather the d^2 local
//    blocks inside the slice, factor the dense block redundantly, keep the own cyclic part (zeros below the
//    global diagonal -- the slots the reference's benchmarked NoReplication policy leaves stale at P > 1).
//  * CholeskyQR2 1D (c == 1): local Gram (split-K DMMA), one all-reduce of n x n, replicated potrf+trtri, local apply
//    (cacqr.hpp:5-29,172-193; cacqr/policy.h:78-85).
//
// NCCL is dlopen'ed ("libnccl.so.2": the copy torch already mapped when the caller is a torch process, else the
// system one) so that the library has no link-time dependency and never mixes two NCCL builds in one process.
#include "dist.cuh"
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  return api;
}

bool nccl_load(std::string* why) {
  NcclApi& a = nccl();
  if (a.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) { *why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
#define LD(field, sym)                                                   \
  *(void**)(&a.field) = dlsym(a.lib, sym);                               \
  if (!a.field) { *why = std::string("missing NCCL symbol ") + sym; a.lib = nullptr; return false; }
  LD(GetUniqueId, "ncclGetUniqueId"); LD(CommInitRank, "ncclCommInitRank"); LD(CommSplit, "ncclCommSplit");
  LD(CommDestroy, "ncclCommDestroy"); LD(AllReduce, "ncclAllReduce"); LD(AllGather, "ncclAllGather");
  LD(Broadcast, "ncclBroadcast"); LD(Send, "ncclSend"); LD(Recv, "ncclRecv"); LD(GroupStart, "ncclGroupStart");
  LD(GroupEnd, "ncclGroupEnd"); LD(GetErrorString, "ncclGetErrorString");
#undef LD
  return true;
}

#define CAP_NCCL(call)                                                                                              \
  do {                                                                                                               \
    int r__ = (call);                                                                                                \
    if (r__ != ncclSuccess) {                                                                                        \
      ctx->set_error(std::string(#call) + ": " + nccl().GetErrorString(r__) + " (" + __FILE__ + ":" +                \
                     std::to_string(__LINE__) + ")");                                                                \
      return CAPITAL_ERR_COMM;                                                                                       \
    }                                                                                                                \
  } while (0)

inline int rank_of(const capital_grid_t& g, int x, int y, int z) { return y * g.c * g.d + x * g.c + z; }  // topology.h:81-83 inverted

// ---- small kernels used only by the distributed schedules ---------------------------------------------------
// C = beta * C + P on an s_m x s_n block (upper_only: local i <= j entries only)
__global__ void axpby_kernel(long long rows, long long cols, const double* __restrict__ P, long long ldp, double beta, double* __restrict__ C,
                             long long ldc, int upper_only) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / rows, r = idx - c * rows;
    if (upper_only && r > c) continue;
    const double v = P[c * ldp + r];
    C[c * ldc + r] = beta == 0.0 ? v : beta * C[c * ldc + r] + v;
  }
}
__global__ void axpby_off_kernel(long long rows, long long cols, const double* __restrict__ P, long long ldp, double beta, double* __restrict__ C,
                                 long long ldc, int upper_only, long long col0) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / rows, r = idx - c * rows;
    if (upper_only && r > c + col0) continue;
    const double v = P[c * ldp + r];
    C[c * ldc + r] = beta == 0.0 ? v : beta * C[c * ldc + r] + v;
  }
}
// gathered[(x' + d y')] = local block (s x s, ld lds) of slice rank x' + d y'  ->  dense (s d) x (s d) block, upper part
// (util::block_to_cyclic_*, util.hpp:56-133)
__global__ void blocks_to_dense_kernel(int s, int d, const double* __restrict__ gathered, long long lds, double* __restrict__ dense,
                                       long long ldd) {
  const long long b = (long long)s * d, total = b * b;
  const long long blk = lds * s;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long gx = idx / b, gy = idx - gx * b;  // col, row
    double v = 0.0;
    if (gy <= gx) {
      const int xo = (int)(gx % d), yo = (int)(gy % d);
      v = gathered[(xo + (long long)d * yo) * blk + (gx / d) * lds + (gy / d)];
    }
    dense[gx * ldd + gy] = v;
  }
}
// own cyclic part of a dense block: loc(j, i) = dense(y + d j, x + d i)   (util::cyclic_to_local, util.hpp:135-164)
// transposed != 0: loc(j, i) = dense(x' ...) of the TRANSPOSED dense matrix, i.e. dense(x + d i, y + d j)^T handled by caller
__global__ void dense_to_local_kernel(int s, int d, int x, int y, const double* __restrict__ dense, long long ldd, double* __restrict__ loc,
                                      long long ldl) {
  const long long total = (long long)s * s;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / s, j = idx - i * s;
    loc[i * ldl + j] = dense[(x + (long long)d * i) * ldd + (y + (long long)d * j)];
  }
}

inline int grid_for(const capital_ctx* ctx, long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

struct Dist {
  capital_ctx* ctx;
  cudaStream_t st;
  const capital_grid_t& g;
  int64_t L, ld;
  double *W, *R, *Ri, *RiT;
  int64_t bc_local;
  int split;
  // contiguous transfer buffers
  double *bufX, *bufY, *bufP, *bufS;  // fetched X, fetched Y, partial/all-reduced product, send staging (2 blocks)
  ncclComm_t world, depth, slice;
  // host-pointer callers: finished column ranges are packed and copied out while the rest of the factorization runs
  bool stream_out = false, rinv_streams = false;
  double *dR = nullptr, *dRinv = nullptr, *hR = nullptr, *hRinv = nullptr;
  int64_t cols_out = 0, rinv_cols_out = 0;
  cudaEvent_t e_out = nullptr;
};

// pack a (rows x cols) window into a contiguous buffer with even leading dimension
inline int64_t packed_ld(int64_t rows) { return round_up(rows, 2); }

capital_status_t product_pipelined(Dist& D, int64_t m, int64_t n, int64_t k, double alpha, const double* X, int64_t ldx, const double* Y,
                                   int64_t ldy, double beta, double* C, int64_t ldc, int flags, int nch);

// One distributed product  C <- beta*C + alpha * X^T Y  (all matrices are windows of cyclically distributed globals;
// local windows: X: k x m, Y: k x n, C: m x n).  X and Y are the LOCAL windows of this rank; the blocks actually
// multiplied are the ones owned by (y,z,z) and (x,z,z).
capital_status_t product(Dist& D, int64_t m, int64_t n, int64_t k, double alpha, const double* X, int64_t ldx, const double* Y,
                         int64_t ldy, double beta, double* C, int64_t ldc, int flags) {
  capital_ctx* ctx = D.ctx;
  const capital_grid_t& g = D.g;
  const int d = g.d, c = g.c, me = g.rank;
  if (ctx->dist_pipeline && ctx->comm_stream && n >= 2048 && m >= 1024 && k >= 1024)
    return product_pipelined(D, m, n, k, alpha, X, ldx, Y, ldy, beta, C, ldc, flags, 4);
  // The contraction index splits into d owner classes (k mod d = kb); the c layers share them: layer z takes the classes
  // kb = z (mod c) when c <= d (the reference has c == d: exactly one class per layer, summa.hpp:185-193), and when
  // c > d (2 x 1 x 1) the single class is cut into c/d row chunks of the local window.
  const int nslices = c > d ? c : d;
  const int nchunk = c > d ? c / d : 1;
  const int64_t ldp = packed_ld(m);
  bool first = true;
  if (flags & CAPITAL_GEMM_C_UPPER) CAP_CUDA(cudaMemsetAsync(D.bufP, 0, (size_t)ldp * n * 8, D.st));
  for (int sl = g.z % (c < nslices ? c : nslices); sl < nslices; sl += c) {
    const int kb = sl % d, chunk = sl / d;
    int64_t r0 = 0, r1 = k;
    int fl = flags;
    if (nchunk > 1) {
      r0 = (k * chunk / nchunk) & ~(int64_t)1;
      r1 = chunk + 1 == nchunk ? k : ((k * (chunk + 1) / nchunk) & ~(int64_t)1);
      fl &= CAPITAL_GEMM_C_UPPER;  // a row chunk is not aligned with the operand's diagonal any more
    }
    const int64_t kk = r1 - r0;
    if (kk <= 0) continue;
    const int64_t ldk = packed_ld(kk);
    const int srcX = rank_of(g, g.y, kb, g.z);  // owner of X rows = kb, cols = y
    const int srcY = rank_of(g, g.x, kb, g.z);  // owner of Y rows = kb, cols = x
    const bool iAmSrc = (g.y == kb);            // my block is the X block of row y' = x and the Y block of column x
    const double* Xuse = X + r0; int64_t ldxu = ldx;
    const double* Yuse = Y + r0; int64_t ldyu = ldy;
    double* sendX = D.bufS;
    double* sendY = D.bufS + ldk * (m > n ? m : n);
    bool needSendX = false, needSendY = false;
    if (iAmSrc) for (int xx = 0; xx < d; xx++) if (rank_of(g, xx, g.x, g.z) != me) needSendX = true;
    if (iAmSrc) for (int yy = 0; yy < d; yy++) if (rank_of(g, g.x, yy, g.z) != me) needSendY = true;
    if (needSendX) CAP_TRY(copy_block(ctx, D.st, kk, m, X + r0, ldx, sendX, ldk));
    if (needSendY) CAP_TRY(copy_block(ctx, D.st, kk, n, Y + r0, ldy, sendY, ldk));
    if (needSendX || needSendY || srcX != me || srcY != me) {
      CAP_NCCL(nccl().GroupStart());
      if (iAmSrc)  // X destinations: all (xx, y' = my x, z);  Y destinations: all (x' = my x, yy, z)
        for (int t = 0; t < d; t++) {
          const int dx = rank_of(g, t, g.x, g.z), dy = rank_of(g, g.x, t, g.z);
          if (dx != me) CAP_NCCL(nccl().Send(sendX, (size_t)ldk * m, ncclFloat64, dx, D.world, D.st));
          if (dy != me) CAP_NCCL(nccl().Send(sendY, (size_t)ldk * n, ncclFloat64, dy, D.world, D.st));
        }
      if (srcX != me) { CAP_NCCL(nccl().Recv(D.bufX, (size_t)ldk * m, ncclFloat64, srcX, D.world, D.st)); Xuse = D.bufX; ldxu = ldk; }
      if (srcY != me) { CAP_NCCL(nccl().Recv(D.bufY, (size_t)ldk * n, ncclFloat64, srcY, D.world, D.st)); Yuse = D.bufY; ldyu = ldk; }
      CAP_NCCL(nccl().GroupEnd());
    }
    CAP_TRY(gemm_tn(ctx, D.st, m, n, kk, alpha, Xuse, ldxu, Yuse, ldyu, first ? 0.0 : 1.0, D.bufP, ldp, fl));
    first = false;
  }
  if (first) CAP_CUDA(cudaMemsetAsync(D.bufP, 0, (size_t)ldp * n * 8, D.st));  // this layer had no slice
  if (c > 1) CAP_NCCL(nccl().AllReduce(D.bufP, D.bufP, (size_t)ldp * n, ncclFloat64, ncclSum, D.depth, D.st));
  axpby_kernel<<<grid_for(ctx, m * n), 256, 0, D.st>>>(m, n, D.bufP, ldp, beta, C, ldc, (flags & CAPITAL_GEMM_C_UPPER) ? 1 : 0);
  ctx->counters.kernel_launches++;
  CAP_CUDA(cudaGetLastError());
  return CAPITAL_OK;
}

capital_status_t comm_event(capital_ctx* ctx, cudaEvent_t* e) {
  if (ctx->comm_used == ctx->comm_pool.size()) {
    cudaEvent_t ev;
    CAP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->comm_pool.push_back(ev);
  }
  *e = ctx->comm_pool[ctx->comm_used++];
  return CAPITAL_OK;
}

// Same product, software-pipelined over column chunks of the output so that the transfers hide behind the DMMA work
// (the reference issues Bcast / GEMM / Allreduce strictly one after the other; its num_chunks option only chunks the
// communication, summa.hpp:195-249).  Two streams: NCCL traffic on ctx->comm_stream, compute on D.st, joined by events:
//     comm:     X | Y_0 | Y_1 | ... | Y_{c-1} |        AR_0 | AR_1 | ...
//     compute:       wait X,Y_0: G_0 | wait Y_1: G_1 | ...        wait AR_0: C_0 | ...
// Only the fetch of X and Y_0 and the last all-reduce stay exposed.
capital_status_t product_pipelined(Dist& D, int64_t m, int64_t n, int64_t k, double alpha, const double* X, int64_t ldx, const double* Y,
                                   int64_t ldy, double beta, double* C, int64_t ldc, int flags, int nch) {
  capital_ctx* ctx = D.ctx;
  const capital_grid_t& g = D.g;
  const int d = g.d, c = g.c, me = g.rank;
  cudaStream_t CS = D.st, NS = ctx->comm_stream;
  ctx->comm_used = 0;
  const int nslices = c > d ? c : d;
  const int nchunk = c > d ? c / d : 1;
  const int64_t ldp = packed_ld(m);
  const int64_t cw = round_up(ceil_div(n, nch), 64);  // chunk width (columns)
  const int nc_eff = (int)ceil_div(n, cw);
  std::vector<cudaEvent_t> e_g(nc_eff, nullptr);
  if (flags & CAPITAL_GEMM_C_UPPER) CAP_CUDA(cudaMemsetAsync(D.bufP, 0, (size_t)ldp * n * 8, CS));
  bool first = true;
  int last_slice = -1;
  for (int sl = g.z; sl < nslices; sl += c) last_slice = sl;
  for (int sl = g.z; sl < nslices; sl += c) {
    const int kb = sl % d, chunk = sl / d;
    int64_t r0 = 0, r1 = k;
    int fl = flags;
    if (nchunk > 1) {
      r0 = (k * chunk / nchunk) & ~(int64_t)1;
      r1 = chunk + 1 == nchunk ? k : ((k * (chunk + 1) / nchunk) & ~(int64_t)1);
      fl &= CAPITAL_GEMM_C_UPPER;
    }
    const int64_t kk = r1 - r0;
    if (kk <= 0) continue;
    const int64_t ldk = packed_ld(kk);
    const int srcX = rank_of(g, g.y, kb, g.z), srcY = rank_of(g, g.x, kb, g.z);
    const bool iAmSrc = (g.y == kb);
    const double* Xuse = X + r0; int64_t ldxu = ldx;
    const double* Yuse = Y + r0; int64_t ldyu = ldy;
    double* sendX = D.bufS;
    double* sendY = D.bufS + ldk * (m > n ? m : n);
    bool needSendX = false, needSendY = false;
    if (iAmSrc) for (int t = 0; t < d; t++) { if (rank_of(g, t, g.x, g.z) != me) needSendX = true; if (rank_of(g, g.x, t, g.z) != me) needSendY = true; }
    if (needSendX) CAP_TRY(copy_block(ctx, CS, kk, m, X + r0, ldx, sendX, ldk));
    if (needSendY) CAP_TRY(copy_block(ctx, CS, kk, n, Y + r0, ldy, sendY, ldk));
    // comm stream may start once the staging copies are done and the previous users of bufX / bufY (earlier GEMMs on CS) are finished
    cudaEvent_t e_pack, e_x;
    CAP_TRY(comm_event(ctx, &e_pack));
    CAP_CUDA(cudaEventRecord(e_pack, CS));
    CAP_CUDA(cudaStreamWaitEvent(NS, e_pack, 0));
    const bool anyX = needSendX || srcX != me, anyY = needSendY || srcY != me;
    if (anyX) {
      CAP_NCCL(nccl().GroupStart());
      if (iAmSrc) for (int t = 0; t < d; t++) { const int dx = rank_of(g, t, g.x, g.z); if (dx != me) CAP_NCCL(nccl().Send(sendX, (size_t)ldk * m, ncclFloat64, dx, D.world, NS)); }
      if (srcX != me) { CAP_NCCL(nccl().Recv(D.bufX, (size_t)ldk * m, ncclFloat64, srcX, D.world, NS)); Xuse = D.bufX; ldxu = ldk; }
      CAP_NCCL(nccl().GroupEnd());
    }
    CAP_TRY(comm_event(ctx, &e_x));
    CAP_CUDA(cudaEventRecord(e_x, NS));
    if (srcY != me) { Yuse = D.bufY; ldyu = ldk; }
    std::vector<cudaEvent_t> e_y(nc_eff, nullptr);
    for (int j = 0; j < nc_eff; j++) {
      const int64_t c0 = (int64_t)j * cw, nc = (c0 + cw <= n) ? cw : n - c0;
      if (anyY) {
        CAP_NCCL(nccl().GroupStart());
        if (iAmSrc) for (int t = 0; t < d; t++) { const int dy = rank_of(g, g.x, t, g.z); if (dy != me) CAP_NCCL(nccl().Send(sendY + c0 * ldk, (size_t)ldk * nc, ncclFloat64, dy, D.world, NS)); }
        if (srcY != me) CAP_NCCL(nccl().Recv(D.bufY + c0 * ldk, (size_t)ldk * nc, ncclFloat64, srcY, D.world, NS));
        CAP_NCCL(nccl().GroupEnd());
      }
      CAP_TRY(comm_event(ctx, &e_y[j]));
      CAP_CUDA(cudaEventRecord(e_y[j], NS));
    }
    CAP_CUDA(cudaStreamWaitEvent(CS, e_x, 0));
    for (int j = 0; j < nc_eff; j++) {
      const int64_t c0 = (int64_t)j * cw, nc = (c0 + cw <= n) ? cw : n - c0;
      CAP_CUDA(cudaStreamWaitEvent(CS, e_y[j], 0));
      CAP_TRY(gemm_tn_off(ctx, CS, m, nc, kk, alpha, Xuse, ldxu, Yuse + c0 * ldyu, ldyu, first ? 0.0 : 1.0, D.bufP + c0 * ldp, ldp, fl, 0, (int)c0));
      if (sl == last_slice && c > 1) {
        CAP_TRY(comm_event(ctx, &e_g[j]));
        CAP_CUDA(cudaEventRecord(e_g[j], CS));
      }
    }
    first = false;
  }
  if (first) CAP_CUDA(cudaMemsetAsync(D.bufP, 0, (size_t)ldp * n * 8, CS));
  const int up = (flags & CAPITAL_GEMM_C_UPPER) ? 1 : 0;
  for (int j = 0; j < nc_eff; j++) {
    const int64_t c0 = (int64_t)j * cw, nc = (c0 + cw <= n) ? cw : n - c0;
    if (c > 1) {
      if (e_g[j]) CAP_CUDA(cudaStreamWaitEvent(NS, e_g[j], 0));
      else { cudaEvent_t e; CAP_TRY(comm_event(ctx, &e)); CAP_CUDA(cudaEventRecord(e, CS)); CAP_CUDA(cudaStreamWaitEvent(NS, e, 0)); }
      CAP_NCCL(nccl().AllReduce(D.bufP + c0 * ldp, D.bufP + c0 * ldp, (size_t)ldp * nc, ncclFloat64, ncclSum, D.depth, NS));
      cudaEvent_t e_r;
      CAP_TRY(comm_event(ctx, &e_r));
      CAP_CUDA(cudaEventRecord(e_r, NS));
      CAP_CUDA(cudaStreamWaitEvent(CS, e_r, 0));
    }
    // C chunk = beta * C + P chunk; for upper-only outputs the mask row <= col uses the global column index
    axpby_off_kernel<<<grid_for(ctx, m * nc), 256, 0, CS>>>(m, nc, D.bufP + c0 * ldp, ldp, beta, C + c0 * ldc, ldc, up, c0);
    ctx->counters.kernel_launches++;
    CAP_CUDA(cudaGetLastError());
  }
  return CAPITAL_OK;
}

// global transpose of a local window: dst(local n x m) = [window of the partner (y,x,z)]^T   (util::transpose, util.hpp:232-247,
// followed by the local transpose the reference defers to its BLAS flags)
capital_status_t transpose_dist(Dist& D, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd) {
  capital_ctx* ctx = D.ctx;
  const capital_grid_t& g = D.g;
  const int partner = rank_of(g, g.y, g.x, g.z);
  if (partner == g.rank) return transpose_block(ctx, D.st, rows, cols, src, lds, dst, ldd, 1.0);
  const int64_t ldr = packed_ld(rows);
  CAP_TRY(copy_block(ctx, D.st, rows, cols, src, lds, D.bufS, ldr));
  CAP_NCCL(nccl().GroupStart());
  CAP_NCCL(nccl().Send(D.bufS, (size_t)ldr * cols, ncclFloat64, partner, D.world, D.st));
  CAP_NCCL(nccl().Recv(D.bufX, (size_t)ldr * cols, ncclFloat64, partner, D.world, D.st));
  CAP_NCCL(nccl().GroupEnd());
  return transpose_block(ctx, D.st, rows, cols, D.bufX, ldr, dst, ldd, 1.0);
}

// replicate-everything base case on the local window at offset `o` of size s (local); dense size b = s d
capital_status_t base_case(Dist& D, int64_t o, int64_t s) {
  capital_ctx* ctx = D.ctx;
  const capital_grid_t& g = D.g;
  const int d = g.d;
  const int64_t b = s * d, ldb = round_up(b, 16), lds = packed_ld(s);
  double *gath, *dW, *dR, *dRi, *dRiT;
  CAP_TRY(ctx->workspace("bc_gather", (size_t)lds * s * d * d * 8, (void**)&gath));
  CAP_TRY(ctx->workspace("bc_W", (size_t)ldb * b * 8, (void**)&dW));
  CAP_TRY(ctx->workspace("bc_R", (size_t)ldb * b * 8, (void**)&dR));
  CAP_TRY(ctx->workspace("bc_Ri", (size_t)ldb * b * 8, (void**)&dRi));
  CAP_TRY(ctx->workspace("bc_RiT", (size_t)ldb * b * 8, (void**)&dRiT));
  double* Wo = D.W + o * D.ld + o;
  CAP_TRY(copy_block(ctx, D.st, s, s, Wo, D.ld, D.bufS, lds));
  if (g.size == 1) CAP_CUDA(cudaMemcpyAsync(gath, D.bufS, (size_t)lds * s * 8, cudaMemcpyDeviceToDevice, D.st));
  else CAP_NCCL(nccl().AllGather(D.bufS, gath, (size_t)lds * s, ncclFloat64, D.slice, D.st));  // policy.h:176
  blocks_to_dense_kernel<<<grid_for(ctx, b * b), 256, 0, D.st>>>((int)s, d, gath, lds, dW, ldb);
  ctx->counters.kernel_launches++;
  CAP_CUDA(cudaGetLastError());
  CAP_CUDA(cudaMemsetAsync(dRi, 0, (size_t)ldb * b * 8, D.st));
  CAP_CUDA(cudaMemsetAsync(dRiT, 0, (size_t)ldb * b * 8, D.st));
  CAP_CUDA(cudaMemsetAsync(dR, 0, (size_t)ldb * b * 8, D.st));
  CAP_TRY(cholinv_local(ctx, D.st, b, dW, ldb, dR, ldb, dRi, ldb, dRiT, ldb, true, b, 1));  // potrf + trtri, policy.h:199-201
  const int gr = grid_for(ctx, s * s);
  dense_to_local_kernel<<<gr, 256, 0, D.st>>>((int)s, d, g.x, g.y, dR, ldb, D.R + o * D.ld + o, D.ld);
  dense_to_local_kernel<<<gr, 256, 0, D.st>>>((int)s, d, g.x, g.y, dRi, ldb, D.Ri + o * D.ld + o, D.ld);
  dense_to_local_kernel<<<gr, 256, 0, D.st>>>((int)s, d, g.x, g.y, dRiT, ldb, D.RiT + o * D.ld + o, D.ld);
  ctx->counters.kernel_launches += 3;
  CAP_CUDA(cudaGetLastError());
  return CAPITAL_OK;
}

capital_status_t dist_io_event(capital_ctx* ctx, cudaEvent_t* e) {
  if (ctx->io_used == ctx->io_pool.size()) {
    cudaEvent_t ev;
    CAP_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->io_pool.push_back(ev);
  }
  *e = ctx->io_pool[ctx->io_used++];
  return CAPITAL_OK;
}
// local columns [cols_out, col_end) of R are final (of Rinv too left of the top split, and up to the next split when the top-level
// inverse block is skipped): pack that contiguous range of the packed triangle and start its D2H on the copy-out stream
capital_status_t dist_left_done(Dist& D, int64_t col_end, int depth) {
  capital_ctx* ctx = D.ctx;
  const int64_t c0 = D.cols_out;
  if (col_end <= c0) return CAPITAL_OK;
  const bool rinv_too = depth == 0 || (depth == 1 && D.rinv_streams && D.rinv_cols_out == c0);
  const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = (size_t)col_end * (col_end + 1) / 2 - off;
  CAP_TRY(pack_upper(ctx, D.st, D.L, D.R, D.ld, D.dR, 0, c0, col_end));
  if (rinv_too) CAP_TRY(pack_upper(ctx, D.st, D.L, D.Ri, D.ld, D.dRinv, 0, c0, col_end));
  D.cols_out = col_end;
  if (rinv_too) D.rinv_cols_out = col_end;
  cudaEvent_t e;
  CAP_TRY(dist_io_event(ctx, &e));
  CAP_CUDA(cudaEventRecord(e, D.st));
  CAP_CUDA(cudaStreamWaitEvent(ctx->copy_out, e, 0));
  if (D.hR) { CAP_CUDA(cudaMemcpyAsync(D.hR + off, D.dR + off, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_out)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
  if (D.hRinv && rinv_too) { CAP_CUDA(cudaMemcpyAsync(D.hRinv + off, D.dRinv + off, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_out)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
  CAP_TRY(dist_io_event(ctx, &D.e_out));
  CAP_CUDA(cudaEventRecord(D.e_out, ctx->copy_out));
  return CAPITAL_OK;
}

// cholinv::invoke (cholinv.hpp:87-165) on the local window [o, o+s)
capital_status_t invoke(Dist& D, int64_t o, int64_t s, bool complete, int depth = 0) {
  const int64_t s1 = s >> D.split;
  if (s <= D.bc_local || s1 < D.split || s1 == 0) return base_case(D, o, s);
  const int64_t s2 = s - s1, ld = D.ld;
  double* W12 = D.W + (o + s1) * ld + o;
  double* W21 = D.W + o * ld + (o + s1);
  double* W22 = D.W + (o + s1) * ld + (o + s1);
  double* R12 = D.R + (o + s1) * ld + o;
  double* Ri11 = D.Ri + o * ld + o;
  double* Ri12 = D.Ri + (o + s1) * ld + o;
  double* Ri22 = D.Ri + (o + s1) * ld + (o + s1);
  double* RiT11 = D.RiT + o * ld + o;
  double* RiT21 = D.RiT + o * ld + (o + s1);
  CAP_TRY(invoke(D, o, s1, true, depth + 1));
  if (D.stream_out && depth <= 3 && o + s == D.L) CAP_TRY(dist_left_done(D, o + s1, depth));  // right spine
  CAP_TRY(product(D, s1, s2, s1, 1.0, Ri11, ld, W12, ld, 0.0, R12, ld, CAPITAL_GEMM_A_UPPER));       // cholinv.hpp:116-122
  CAP_TRY(product(D, s2, s2, s1, -1.0, R12, ld, R12, ld, 1.0, W22, ld, CAPITAL_GEMM_C_UPPER));        // :131-134
  CAP_TRY(invoke(D, o + s1, s2, true, depth + 1));
  if (complete) {                                                                                       // :147-155
    CAP_TRY(product(D, s2, s1, s1, 1.0, R12, ld, RiT11, ld, 0.0, W21, ld, CAPITAL_GEMM_B_LOWER));
    CAP_TRY(product(D, s1, s2, s2, -1.0, W21, ld, Ri22, ld, 0.0, Ri12, ld, CAPITAL_GEMM_B_UPPER));
    CAP_TRY(transpose_dist(D, s1, s2, Ri12, ld, RiT21, ld));
  }
  return CAPITAL_OK;
}

capital_status_t need_comm(capital_ctx* ctx) {
  if (ctx->grid.size > 1 && !ctx->comm_world) {
    ctx->set_error("multi-GPU grid but capital_comm_init was not called");
    return CAPITAL_ERR_COMM;
  }
  return CAPITAL_OK;
}

capital_status_t allreduce_scalars(capital_ctx* ctx, double* dptr, int count) {
  if (ctx->grid.size > 1) CAP_NCCL(nccl().AllReduce(dptr, dptr, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm_world, ctx->stream));
  return CAPITAL_OK;
}

}  // namespace

extern "C" capital_status_t capital_comm_unique_id(void* out128) {
  std::string why;
  if (!out128 || !nccl_load(&why)) return CAPITAL_ERR_COMM;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return CAPITAL_ERR_COMM;
  memcpy(out128, &id, 128);
  return CAPITAL_OK;
}

extern "C" capital_status_t capital_comm_init(capital_ctx* ctx, const void* uid) {
  if (!ctx || !uid) return CAPITAL_ERR_INVALID;
  CAP_CUDA(cudaSetDevice(ctx->device));
  std::string why;
  if (!nccl_load(&why)) { ctx->set_error(why); return CAPITAL_ERR_COMM; }
  const capital_grid_t& g = ctx->grid;
  ncclUniqueId id;
  memcpy(&id, uid, 128);
  ncclComm_t world = nullptr, depth = nullptr, slice = nullptr;
  CAP_NCCL(nccl().CommInitRank(&world, g.size, id, g.rank));
  ctx->comm_world = world;
  // sub-communicators of topo::square (topology.h:84-94): depth = same (x,y), slice = same z
  CAP_NCCL(nccl().CommSplit(world, g.y * g.d + g.x, g.z, &depth, nullptr));
  CAP_NCCL(nccl().CommSplit(world, g.z, g.y * g.d + g.x, &slice, nullptr));
  ctx->comm_depth = depth;
  ctx->comm_slice = slice;
  CAP_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("CAPITAL_DIST_PIPELINE")) ctx->dist_pipeline = atoi(e) != 0;
  return CAPITAL_OK;
}

void dist_destroy(capital_ctx* ctx) {
  if (!nccl().lib) return;
  if (ctx->comm_depth) nccl().CommDestroy((ncclComm_t)ctx->comm_depth);
  if (ctx->comm_slice) nccl().CommDestroy((ncclComm_t)ctx->comm_slice);
  if (ctx->comm_world) nccl().CommDestroy((ncclComm_t)ctx->comm_world);
  ctx->comm_depth = ctx->comm_slice = ctx->comm_world = nullptr;
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
  for (cudaEvent_t e : ctx->comm_pool) cudaEventDestroy(e);
  ctx->comm_pool.clear();
}

capital_status_t dist_cholinv_factor(capital_ctx* ctx, const double* A_local, int64_t n, const capital_cholinv_args_t* args,
                                     capital_structure_t ostruct, double* R_local, double* Rinv_local) {
  const capital_grid_t& g = ctx->grid;
  CAP_TRY(need_comm(ctx));
  // the reference requires c == d (summa.hpp:16-31); c | d and d | c grids (2x1x1, 1x2x2) are this library's extension
  if ((g.c % g.d != 0 && g.d % g.c != 0) || n % g.d != 0) {
    ctx->set_error("distributed cholinv needs a grid with c | d or d | c, and d | n");
    return CAPITAL_ERR_UNSUPPORTED;
  }
  const int64_t L = n / g.d, ld = round_up(L, 16);
  const size_t out_count = ostruct == CAPITAL_UPPERTRI_PACKED ? (size_t)L * (L + 1) / 2 : (size_t)L * L;
  cudaStream_t st = ctx->stream;
  CAP_CUDA(cudaEventRecord(ctx->ev_start, st));
  Dist D{ctx, st, g};
  D.L = L; D.ld = ld; D.split = (int)args->split;
  D.bc_local = capital_cholinv_bc_dimension(L, g.c, g.d, args->bc_mult_dim) / g.d;
  D.world = (ncclComm_t)ctx->comm_world; D.depth = (ncclComm_t)ctx->comm_depth; D.slice = (ncclComm_t)ctx->comm_slice;
  double *dR, *dRinv;
  CAP_TRY(ctx->workspace("W", (size_t)ld * L * 8, (void**)&D.W));
  CAP_TRY(ctx->workspace("Rm", (size_t)ld * L * 8, (void**)&D.R));
  CAP_TRY(ctx->workspace("Ri", (size_t)ld * L * 8, (void**)&D.Ri));
  CAP_TRY(ctx->workspace("RiT", (size_t)ld * L * 8, (void**)&D.RiT));
  const int64_t half = L - (L >> D.split) > (L >> D.split) ? L - (L >> D.split) : (L >> D.split);
  const size_t blk = (size_t)packed_ld(half) * half * 8 + 4096;
  CAP_TRY(ctx->workspace("xferX", blk, (void**)&D.bufX));
  CAP_TRY(ctx->workspace("xferY", blk, (void**)&D.bufY));
  CAP_TRY(ctx->workspace("xferP", blk, (void**)&D.bufP));
  CAP_TRY(ctx->workspace("xferS", 2 * blk, (void**)&D.bufS));
  CAP_TRY(cap_stage_out_begin(ctx, R_local, out_count, "R_out", &dR));
  CAP_TRY(cap_stage_out_begin(ctx, Rinv_local, out_count, "Rinv_out", &dRinv));
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), st));
  CAP_CUDA(cudaMemsetAsync(D.Ri, 0, (size_t)ld * L * 8, st));
  CAP_CUDA(cudaMemsetAsync(D.RiT, 0, (size_t)ld * L * 8, st));
  CAP_CUDA(cudaMemsetAsync(D.R, 0, (size_t)ld * L * 8, st));
  if (cap_is_device_ptr(A_local)) {
    CAP_TRY(copy_block(ctx, st, L, L, A_local, L, D.W, ld));  // serialize(A -> R), cholinv.hpp:13
  } else {
    // host caller: only the (local) upper triangle is read, so only rows [0, column chunk end) travel
    const int64_t chunk = round_up(ceil_div(L, 16), 64);
    for (int64_t c0 = 0; c0 < L; c0 += chunk) {
      const int64_t nc = (c0 + chunk <= L) ? chunk : L - c0, rows = c0 + nc;
      CAP_CUDA(cudaMemcpy2DAsync(D.W + c0 * ld, (size_t)ld * 8, A_local + c0 * L, (size_t)L * 8, (size_t)rows * 8, (size_t)nc,
                                 cudaMemcpyHostToDevice, st));
      ctx->counters.h2d_bytes += rows * nc * 8;
    }
  }
  D.dR = dR; D.dRinv = dRinv;
  D.hR = dR != R_local ? R_local : nullptr;
  D.hRinv = dRinv != Rinv_local ? Rinv_local : nullptr;
  D.stream_out = ostruct == CAPITAL_UPPERTRI_PACKED && (D.hR || D.hRinv) && L >= 2048;
  D.rinv_streams = args->complete_inv == 0;
  ctx->io_used = 0;
  CAP_TRY(invoke(D, 0, L, args->complete_inv != 0));
  if (ostruct == CAPITAL_UPPERTRI_PACKED) {
    {
      const int64_t c0 = D.cols_out;
      const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = out_count - off;
      CAP_TRY(pack_upper(ctx, st, L, D.R, ld, dR, 0, c0, L));
      if (D.hR) { CAP_CUDA(cudaMemcpyAsync(D.hR + off, dR + off, cnt * 8, cudaMemcpyDeviceToHost, st)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
    }
    {
      const int64_t c0 = D.rinv_cols_out;
      const size_t off = (size_t)c0 * (c0 + 1) / 2, cnt = out_count - off;
      CAP_TRY(pack_upper(ctx, st, L, D.Ri, ld, dRinv, 0, c0, L));
      if (D.hRinv) { CAP_CUDA(cudaMemcpyAsync(D.hRinv + off, dRinv + off, cnt * 8, cudaMemcpyDeviceToHost, st)); ctx->counters.d2h_bytes += (int64_t)cnt * 8; }
    }
    if (D.e_out) CAP_CUDA(cudaStreamWaitEvent(st, D.e_out, 0));
  } else {
    CAP_TRY(triu_copy(ctx, st, L, D.R, ld, dR, L, 0));
    CAP_TRY(triu_copy(ctx, st, L, D.Ri, ld, dRinv, L, 0));
    CAP_TRY(cap_stage_out_end(ctx, R_local, out_count, dR));
    CAP_TRY(cap_stage_out_end(ctx, Rinv_local, out_count, dRinv));
  }
  CAP_CUDA(cudaEventRecord(ctx->ev_stop, st));
  return cap_check_info(ctx);
}

capital_status_t dist_cholinv_residual(capital_ctx* ctx, const double* A_local, int64_t n, capital_structure_t structure,
                                       const double* R_local, double* residual) {
  const capital_grid_t& g = ctx->grid;
  CAP_TRY(need_comm(ctx));
  if ((g.c % g.d != 0 && g.d % g.c != 0) || n % g.d != 0) return CAPITAL_ERR_UNSUPPORTED;
  const int64_t L = n / g.d, ld = round_up(L, 16);
  cudaStream_t st = ctx->stream;
  const size_t r_count = structure == CAPITAL_UPPERTRI_PACKED ? (size_t)L * (L + 1) / 2 : (size_t)L * L;
  const double *dA, *dRin;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)L * L, "A_in", &dA));
  CAP_TRY(cap_stage_in(ctx, R_local, r_count, "R_in", &dRin));
  Dist D{ctx, st, g};
  D.L = L; D.ld = ld; D.split = 1; D.bc_local = L;
  D.world = (ncclComm_t)ctx->comm_world; D.depth = (ncclComm_t)ctx->comm_depth; D.slice = (ncclComm_t)ctx->comm_slice;
  double *E, *Rr;
  CAP_TRY(ctx->workspace("W", (size_t)ld * L * 8, (void**)&E));
  CAP_TRY(ctx->workspace("Rm", (size_t)ld * L * 8, (void**)&Rr));
  const size_t blk = (size_t)packed_ld(L) * L * 8 + 4096;
  CAP_TRY(ctx->workspace("xferX", blk, (void**)&D.bufX));
  CAP_TRY(ctx->workspace("xferY", blk, (void**)&D.bufY));
  CAP_TRY(ctx->workspace("xferP", blk, (void**)&D.bufP));
  CAP_TRY(ctx->workspace("xferS", 2 * blk, (void**)&D.bufS));
  if (structure == CAPITAL_UPPERTRI_PACKED) CAP_TRY(unpack_upper(ctx, st, L, dRin, Rr, ld));
  else CAP_TRY(triu_copy(ctx, st, L, dRin, L, Rr, ld, 0));
  if (g.y > g.x) {  // util::remove_triangle (validate.hpp:11): local diagonal is below the global diagonal there
    CAP_TRY(triu_copy(ctx, st, L, Rr, ld, Rr, ld, 1));
  }
  CAP_TRY(copy_block(ctx, st, L, L, dA, L, E, ld));
  CAP_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, 2 * sizeof(double), st));
  CAP_TRY(sumsq_block(ctx, st, L, L, E, ld, 1, g.x, g.y, g.d, ctx->d_scalars + 1));
  // E = R^T R - A  (validate.hpp:35).  No C_UPPER: on ranks with y > x the local diagonal is outside the global upper part anyway.
  CAP_TRY(product(D, L, L, L, 1.0, Rr, ld, Rr, ld, -1.0, E, ld, CAPITAL_GEMM_A_UPPER | CAPITAL_GEMM_B_UPPER));
  CAP_TRY(sumsq_block(ctx, st, L, L, E, ld, 1, g.x, g.y, g.d, ctx->d_scalars));
  CAP_TRY(allreduce_scalars(ctx, ctx->d_scalars, 2));
  double h[2];
  CAP_CUDA(cudaMemcpyAsync(h, ctx->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CAP_CUDA(cudaStreamSynchronize(st));
  *residual = sqrt(h[0]) / sqrt(h[1]);
  return CAPITAL_OK;
}

capital_status_t dist_summa_gemm_tn(capital_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A_local,
                                    const double* B_local, double beta, double* C_local) {
  const capital_grid_t& g = ctx->grid;
  CAP_TRY(need_comm(ctx));
  if ((g.c % g.d != 0 && g.d % g.c != 0) || m % g.d || n % g.d || k % g.d) {
    ctx->set_error("summa gemm: needs a grid with c | d or d | c, and d | m, n, k");
    return CAPITAL_ERR_UNSUPPORTED;
  }
  const int64_t ml = m / g.d, nl = n / g.d, kl = k / g.d;
  cudaStream_t st = ctx->stream;
  const double *dA, *dB;
  double* dC;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)kl * ml, "summa_A", &dA));
  CAP_TRY(cap_stage_in(ctx, B_local, (size_t)kl * nl, "summa_B", &dB));
  const bool c_host = !cap_is_device_ptr(C_local);
  if (c_host) {
    const double* tmp;
    CAP_TRY(cap_stage_in(ctx, C_local, (size_t)ml * nl, "summa_C", &tmp));
    dC = const_cast<double*>(tmp);
  } else dC = C_local;
  // TMA needs even leading dimensions: repack operands whose local row count is odd
  const int64_t ldk = packed_ld(kl);
  double *pA = const_cast<double*>(dA), *pB = const_cast<double*>(dB);
  if (ldk != kl) {
    CAP_TRY(ctx->workspace("summa_pA", (size_t)ldk * ml * 8, (void**)&pA));
    CAP_TRY(ctx->workspace("summa_pB", (size_t)ldk * nl * 8, (void**)&pB));
    CAP_TRY(copy_block(ctx, st, kl, ml, dA, kl, pA, ldk));
    CAP_TRY(copy_block(ctx, st, kl, nl, dB, kl, pB, ldk));
  }
  if (g.size == 1) {
    CAP_TRY(gemm_tn(ctx, st, ml, nl, kl, alpha, pA, ldk, pB, ldk, beta, dC, ml, 0));
  } else {
    Dist D{ctx, st, g};
    D.L = 0; D.ld = 0; D.split = 1; D.bc_local = 0;
    D.world = (ncclComm_t)ctx->comm_world; D.depth = (ncclComm_t)ctx->comm_depth; D.slice = (ncclComm_t)ctx->comm_slice;
    const int64_t mx = ml > nl ? ml : nl;
    const size_t blk = (size_t)packed_ld(kl > ml ? kl : ml) * mx * 8 + 4096;
    CAP_TRY(ctx->workspace("xferX", blk, (void**)&D.bufX));
    CAP_TRY(ctx->workspace("xferY", blk, (void**)&D.bufY));
    CAP_TRY(ctx->workspace("xferP", blk, (void**)&D.bufP));
    CAP_TRY(ctx->workspace("xferS", 2 * blk, (void**)&D.bufS));
    CAP_TRY(product(D, ml, nl, kl, alpha, pA, ldk, pB, ldk, beta, dC, ml, 0));
  }
  if (c_host) CAP_TRY(cap_stage_out_end(ctx, C_local, (size_t)ml * nl, dC));
  CAP_CUDA(cudaStreamSynchronize(st));
  return CAPITAL_OK;
}

// ---- CholeskyQR2, 1D --------------------------------------------------------------------------------------------
namespace {
struct Qr {
  capital_ctx* ctx;
  cudaStream_t st;
  int64_t lr, n, ldq, ldn, ldt;
  double *Q, *Qt, *Qt2, *G, *W, *R1, *R2, *Ri, *RiT, *Rt;
};

// sweep_1d (cacqr.hpp:5-29): G = Q^T Q, all-reduce, R = chol(G), Rinv, Q <- Q Rinv.  R lands in `Rout`.
capital_status_t sweep(Qr& q, double* Rout) {
  capital_ctx* ctx = q.ctx;
  cudaStream_t st = q.st;
  const int64_t n = q.n, lr = q.lr;
  CAP_CUDA(cudaMemsetAsync(q.G, 0, (size_t)q.ldn * n * 8, st));
  CAP_TRY(gemm_tn_splitk(ctx, st, n, n, lr, 1.0, q.Q, q.ldq, q.Q, q.ldq, q.G, q.ldn, CAPITAL_GEMM_C_UPPER));  // dsyrk 'U','T' (:15)
  if (ctx->grid.size > 1)
    CAP_NCCL(nccl().AllReduce(q.G, q.G, (size_t)q.ldn * n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm_world, st));  // policy.h:82
  CAP_CUDA(cudaMemsetAsync(q.Ri, 0, (size_t)q.ldn * n * 8, st));
  CAP_CUDA(cudaMemsetAsync(q.RiT, 0, (size_t)q.ldn * n * 8, st));
  CAP_CUDA(cudaMemsetAsync(Rout, 0, (size_t)q.ldn * n * 8, st));
  CAP_TRY(cholinv_local(ctx, st, n, q.G, q.ldn, Rout, q.ldn, q.Ri, q.ldn, q.RiT, q.ldn, true, n, 1));  // potrf + trtri (:20-22)
  // Q <- Q Rinv (dtrmm Right/Upper/NoTrans, :25) as (Q Rinv)^T = Rinv^T Q^T : A = Rinv (upper), B = Q^T
  CAP_TRY(transpose_block(ctx, st, lr, n, q.Q, q.ldq, q.Qt, q.ldt, 1.0));
  CAP_TRY(gemm_tn(ctx, st, n, lr, n, 1.0, q.Ri, q.ldn, q.Qt, q.ldt, 0.0, q.Qt2, q.ldt, CAPITAL_GEMM_A_UPPER));
  CAP_TRY(transpose_block(ctx, st, n, lr, q.Qt2, q.ldt, q.Q, q.ldq, 1.0));
  return CAPITAL_OK;
}
}  // namespace

// ---- CholeskyQR2, 3D grid (c == d) ---------------------------------------------------------------------------------
// qr::cacqr::invoke_3d / sweep_3d (cacqr.hpp:75-120,195-215): Gram matrix by a SUMMA step, cholinv::factor on the n x n Gram
// matrix over the same grid, Q <- Q R^{-1} by a SUMMA trmm.  Here each of those is the distributed A^T B product of this file:
//   G    = Q^T Q                      product(X = Q, Y = Q)                   [row Bcast + dgemm + column Reduce + depth Bcast, :92-99]
//   R, R^{-1} from invoke() on G                                                 [cholinv::factor, :103]
//   Q^T <- R^{-T} Q^T                  product(X = Rinv (upper), Y = Q^T)      [summa trmm Right/Upper, :111]
// with the global transposes done by the partner exchange (util::transpose).  The complete inverse is always formed
// (the reference's block `solve` for complete_inv == 0, :44-73, yields the same Q).
namespace {
struct Qr3 {
  Dist* D;
  int64_t ml, nl, ldq, ldn;
  double *Q, *T1, *T2;
};
capital_status_t sweep3d(Qr3& q, double* Rout) {
  Dist& D = *q.D;
  capital_ctx* ctx = D.ctx;
  const int64_t ml = q.ml, nl = q.nl, ld = q.ldn;
  CAP_TRY(product(D, nl, nl, ml, 1.0, q.Q, q.ldq, q.Q, q.ldq, 0.0, D.W, ld, 0));
  CAP_CUDA(cudaMemsetAsync(D.Ri, 0, (size_t)ld * nl * 8, D.st));
  CAP_CUDA(cudaMemsetAsync(D.RiT, 0, (size_t)ld * nl * 8, D.st));
  CAP_CUDA(cudaMemsetAsync(D.R, 0, (size_t)ld * nl * 8, D.st));
  CAP_TRY(invoke(D, 0, nl, true));
  CAP_TRY(copy_block(ctx, D.st, nl, nl, D.R, ld, Rout, ld));
  CAP_TRY(transpose_dist(D, ml, nl, q.Q, q.ldq, q.T1, ld));                                           // T1 = Q^T
  CAP_TRY(product(D, nl, ml, nl, 1.0, D.Ri, ld, q.T1, ld, 0.0, q.T2, ld, CAPITAL_GEMM_A_UPPER));      // T2 = Rinv^T Q^T
  CAP_TRY(transpose_dist(D, nl, ml, q.T2, ld, q.Q, q.ldq));                                           // Q = T2^T
  return CAPITAL_OK;
}
capital_status_t qr3_setup(capital_ctx* ctx, Dist& D, Qr3& q, int64_t m, int64_t n, const capital_cholinv_args_t* ci_args) {
  const capital_grid_t& g = ctx->grid;
  if (m % g.d || n % g.d) {
    ctx->set_error("cacqr 3D: d must divide m and n");
    return CAPITAL_ERR_UNSUPPORTED;
  }
  q.D = &D;
  q.ml = m / g.d; q.nl = n / g.d; q.ldq = round_up(q.ml, 16); q.ldn = round_up(q.nl, 16);
  D.L = q.nl; D.ld = q.ldn; D.split = ci_args ? (int)ci_args->split : 1;
  if (D.split <= 0) D.split = 1;
  D.bc_local = capital_cholinv_bc_dimension(q.nl, g.c, g.d, ci_args ? ci_args->bc_mult_dim : 0) / g.d;
  D.world = (ncclComm_t)ctx->comm_world; D.depth = (ncclComm_t)ctx->comm_depth; D.slice = (ncclComm_t)ctx->comm_slice;
  const size_t nn = (size_t)q.ldn * q.nl * 8;
  CAP_TRY(ctx->workspace("q3W", nn, (void**)&D.W));
  CAP_TRY(ctx->workspace("q3R", nn, (void**)&D.R));
  CAP_TRY(ctx->workspace("q3Ri", nn, (void**)&D.Ri));
  CAP_TRY(ctx->workspace("q3RiT", nn, (void**)&D.RiT));
  CAP_TRY(ctx->workspace("q3Q", (size_t)q.ldq * q.nl * 8, (void**)&q.Q));
  CAP_TRY(ctx->workspace("q3T1", (size_t)q.ldn * q.ml * 8, (void**)&q.T1));
  CAP_TRY(ctx->workspace("q3T2", (size_t)q.ldn * q.ml * 8, (void**)&q.T2));
  const size_t blk = ((size_t)packed_ld(q.ml) * packed_ld(q.nl) + (size_t)packed_ld(q.nl) * packed_ld(q.nl)) * 8 + 4096;
  CAP_TRY(ctx->workspace("xferX", blk, (void**)&D.bufX));
  CAP_TRY(ctx->workspace("xferY", blk, (void**)&D.bufY));
  CAP_TRY(ctx->workspace("xferP", blk, (void**)&D.bufP));
  CAP_TRY(ctx->workspace("xferS", 2 * blk, (void**)&D.bufS));
  return CAPITAL_OK;
}
capital_status_t cacqr3d_factor(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, int num_iter, const capital_cholinv_args_t* ci_args,
                                capital_structure_t rstruct, double* Q_local, double* R_local) {
  const capital_grid_t& g = ctx->grid;
  cudaStream_t st = ctx->stream;
  CAP_CUDA(cudaEventRecord(ctx->ev_start, st));
  Dist D{ctx, st, g};
  Qr3 q{};
  CAP_TRY(qr3_setup(ctx, D, q, m, n, ci_args));
  const int64_t ml = q.ml, nl = q.nl, ld = q.ldn;
  const double* dA;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)ml * nl, "A_in", &dA));
  const size_t r_count = rstruct == CAPITAL_UPPERTRI_PACKED ? (size_t)nl * (nl + 1) / 2 : (size_t)nl * nl;
  double *dQ, *dR, *R1, *R2, *Rt, *Rf;
  CAP_TRY(cap_stage_out_begin(ctx, Q_local, (size_t)ml * nl, "Q_out", &dQ));
  CAP_TRY(cap_stage_out_begin(ctx, R_local, r_count, "R_out", &dR));
  const size_t nn = (size_t)ld * nl * 8;
  CAP_TRY(ctx->workspace("q3R1", nn, (void**)&R1));
  CAP_TRY(ctx->workspace("q3R2", nn, (void**)&R2));
  CAP_TRY(ctx->workspace("q3Rt", nn, (void**)&Rt));
  CAP_TRY(ctx->workspace("q3Rf", nn, (void**)&Rf));
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), st));
  CAP_TRY(copy_block(ctx, st, ml, nl, dA, ml, q.Q, q.ldq));
  CAP_TRY(sweep3d(q, R1));
  const double* Rfinal = R1;
  if (num_iter > 1) {
    CAP_TRY(sweep3d(q, R2));
    CAP_TRY(transpose_dist(D, nl, nl, R2, ld, Rt, ld));  // R = R2 R1 = (R2^T)^T R1  (cacqr.hpp:207-209)
    CAP_CUDA(cudaMemsetAsync(Rf, 0, nn, st));
    CAP_TRY(product(D, nl, nl, nl, 1.0, Rt, ld, R1, ld, 0.0, Rf, ld, CAPITAL_GEMM_A_LOWER | CAPITAL_GEMM_B_UPPER));
    Rfinal = Rf;
  }
  const int zd = g.y > g.x ? 1 : 0;  // local diagonal is below the global diagonal on those ranks
  if (rstruct == CAPITAL_UPPERTRI_PACKED) CAP_TRY(pack_upper(ctx, st, nl, Rfinal, ld, dR, zd));
  else CAP_TRY(triu_copy(ctx, st, nl, Rfinal, ld, dR, nl, zd));
  CAP_TRY(copy_block(ctx, st, ml, nl, q.Q, q.ldq, dQ, ml));
  CAP_TRY(cap_stage_out_end(ctx, Q_local, (size_t)ml * nl, dQ));
  CAP_TRY(cap_stage_out_end(ctx, R_local, r_count, dR));
  CAP_CUDA(cudaEventRecord(ctx->ev_stop, st));
  return cap_check_info(ctx);
}
capital_status_t cacqr3d_residual(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, const double* Q_local,
                                  capital_structure_t rstruct, const double* R_local, double* residual, double* orthogonality) {
  const capital_grid_t& g = ctx->grid;
  cudaStream_t st = ctx->stream;
  Dist D{ctx, st, g};
  Qr3 q{};
  capital_cholinv_args_t dummy{1, 1, 0, 'U'};
  CAP_TRY(qr3_setup(ctx, D, q, m, n, &dummy));
  const int64_t ml = q.ml, nl = q.nl, ld = q.ldn;
  const size_t r_count = rstruct == CAPITAL_UPPERTRI_PACKED ? (size_t)nl * (nl + 1) / 2 : (size_t)nl * nl;
  const double *dA, *dQ, *dRin;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)ml * nl, "A_in", &dA));
  CAP_TRY(cap_stage_in(ctx, Q_local, (size_t)ml * nl, "Q_in", &dQ));
  CAP_TRY(cap_stage_in(ctx, R_local, r_count, "R_in", &dRin));
  double* Rr = D.R;
  if (rstruct == CAPITAL_UPPERTRI_PACKED) CAP_TRY(unpack_upper(ctx, st, nl, dRin, Rr, ld));
  else CAP_TRY(triu_copy(ctx, st, nl, dRin, nl, Rr, ld, 0));
  if (g.y > g.x) CAP_TRY(triu_copy(ctx, st, nl, Rr, ld, Rr, ld, 1));  // util::remove_triangle (validate.hpp:42)
  CAP_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, 3 * sizeof(double), st));
  // residual: (Q R)^T - A^T = R^T Q^T - A^T
  CAP_TRY(copy_block(ctx, st, ml, nl, dQ, ml, q.Q, q.ldq));
  CAP_TRY(transpose_dist(D, ml, nl, q.Q, q.ldq, q.T1, ld));   // Q^T
  CAP_TRY(copy_block(ctx, st, ml, nl, dA, ml, q.Q, q.ldq));
  CAP_TRY(transpose_dist(D, ml, nl, q.Q, q.ldq, q.T2, ld));   // A^T
  CAP_TRY(sumsq_block(ctx, st, nl, ml, q.T2, ld, 0, 0, 0, 1, ctx->d_scalars + 1));
  CAP_TRY(product(D, nl, ml, nl, 1.0, Rr, ld, q.T1, ld, -1.0, q.T2, ld, CAPITAL_GEMM_A_UPPER));
  CAP_TRY(sumsq_block(ctx, st, nl, ml, q.T2, ld, 0, 0, 0, 1, ctx->d_scalars));
  // orthogonality: Q^T Q - I
  CAP_TRY(copy_block(ctx, st, ml, nl, dQ, ml, q.Q, q.ldq));
  CAP_TRY(product(D, nl, nl, ml, 1.0, q.Q, q.ldq, q.Q, q.ldq, 0.0, D.W, ld, 0));
  if (g.x == g.y) CAP_TRY(sub_identity_local(ctx, st, nl, D.W, ld));
  CAP_TRY(sumsq_block(ctx, st, nl, nl, D.W, ld, 0, 0, 0, 1, ctx->d_scalars + 2));
  CAP_TRY(allreduce_scalars(ctx, ctx->d_scalars, 3));  // every layer holds a replica: all three sums carry the same factor c
  double h[3];
  CAP_CUDA(cudaMemcpyAsync(h, ctx->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CAP_CUDA(cudaStreamSynchronize(st));
  *residual = sqrt(h[0]) / sqrt(h[1]);
  *orthogonality = sqrt(h[2] / g.c) / sqrt((double)n * (double)n);
  return CAPITAL_OK;
}
inline bool use_3d(const capital_grid_t& g) {
  if (g.c != g.d) return false;
  if (g.c > 1) return true;
  const char* e = getenv("CAPITAL_FORCE_QR3D");  // 1x1x1: the reference takes the 1D path (cacqr.hpp:229); tests may force the 3D code
  return e && atoi(e) != 0;
}
}  // namespace

capital_status_t dist_cacqr_factor(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, int num_iter,
                                   const capital_cholinv_args_t* ci_args, capital_structure_t rstruct, double* Q_local, double* R_local) {
  (void)ci_args;
  const capital_grid_t& g = ctx->grid;
  CAP_TRY(need_comm(ctx));
  if (use_3d(g)) return cacqr3d_factor(ctx, A_local, m, n, num_iter, ci_args, rstruct, Q_local, R_local);
  if (g.c != 1) {
    ctx->set_error("cacqr: 1D (c == 1, cacqr.hpp:229) and 3D (c == d, :232) grids are implemented; the tunable c < d grid (:234-246) is not");
    return CAPITAL_ERR_UNSUPPORTED;
  }
  const int64_t lr = ceil_div(m, g.d);
  cudaStream_t st = ctx->stream;
  CAP_CUDA(cudaEventRecord(ctx->ev_start, st));
  Qr q{ctx, st};
  q.lr = lr; q.n = n; q.ldq = round_up(lr, 16); q.ldn = round_up(n, 16); q.ldt = q.ldn;
  const double* dA;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)lr * n, "A_in", &dA));
  const size_t r_count = rstruct == CAPITAL_UPPERTRI_PACKED ? (size_t)n * (n + 1) / 2 : (size_t)n * n;
  double *dQ, *dR;
  CAP_TRY(cap_stage_out_begin(ctx, Q_local, (size_t)lr * n, "Q_out", &dQ));
  CAP_TRY(cap_stage_out_begin(ctx, R_local, r_count, "R_out", &dR));
  CAP_TRY(ctx->workspace("qrQ", (size_t)q.ldq * n * 8, (void**)&q.Q));
  CAP_TRY(ctx->workspace("qrQt", (size_t)q.ldt * lr * 8, (void**)&q.Qt));
  CAP_TRY(ctx->workspace("qrQt2", (size_t)q.ldt * lr * 8, (void**)&q.Qt2));
  const size_t nn = (size_t)q.ldn * n * 8;
  CAP_TRY(ctx->workspace("qrG", nn, (void**)&q.G));
  CAP_TRY(ctx->workspace("qrR1", nn, (void**)&q.R1));
  CAP_TRY(ctx->workspace("qrR2", nn, (void**)&q.R2));
  CAP_TRY(ctx->workspace("qrRi", nn, (void**)&q.Ri));
  CAP_TRY(ctx->workspace("qrRiT", nn, (void**)&q.RiT));
  CAP_TRY(ctx->workspace("qrRt", nn, (void**)&q.Rt));
  CAP_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), st));
  CAP_TRY(copy_block(ctx, st, lr, n, dA, lr, q.Q, q.ldq));  // Q <- A (cacqr.hpp:226)
  CAP_TRY(sweep(q, q.R1));
  const double* Rfinal = q.R1;
  if (num_iter > 1) {
    CAP_TRY(sweep(q, q.R2));
    // R = R2 R1 (dtrmm, cacqr.hpp:185-187) = (R2^T)^T R1 : A = R2^T (lower), B = R1 (upper)
    CAP_TRY(transpose_block(ctx, st, n, n, q.R2, q.ldn, q.Rt, q.ldn, 1.0));
    CAP_TRY(gemm_tn(ctx, st, n, n, n, 1.0, q.Rt, q.ldn, q.R1, q.ldn, 0.0, q.G, q.ldn,
                    CAPITAL_GEMM_A_LOWER | CAPITAL_GEMM_B_UPPER | CAPITAL_GEMM_C_UPPER));
    Rfinal = q.G;
  }
  if (rstruct == CAPITAL_UPPERTRI_PACKED) CAP_TRY(pack_upper(ctx, st, n, Rfinal, q.ldn, dR, 0));
  else CAP_TRY(triu_copy(ctx, st, n, Rfinal, q.ldn, dR, n, 0));
  CAP_TRY(copy_block(ctx, st, lr, n, q.Q, q.ldq, dQ, lr));
  CAP_TRY(cap_stage_out_end(ctx, Q_local, (size_t)lr * n, dQ));
  CAP_TRY(cap_stage_out_end(ctx, R_local, r_count, dR));
  CAP_CUDA(cudaEventRecord(ctx->ev_stop, st));
  return cap_check_info(ctx);
}

capital_status_t dist_cacqr_residual(capital_ctx* ctx, const double* A_local, int64_t m, int64_t n, const double* Q_local,
                                     capital_structure_t rstruct, const double* R_local, double* residual, double* orthogonality) {
  const capital_grid_t& g = ctx->grid;
  CAP_TRY(need_comm(ctx));
  if (use_3d(g)) return cacqr3d_residual(ctx, A_local, m, n, Q_local, rstruct, R_local, residual, orthogonality);
  if (g.c != 1) return CAPITAL_ERR_UNSUPPORTED;
  const int64_t lr = ceil_div(m, g.d);
  cudaStream_t st = ctx->stream;
  const int64_t ldq = round_up(lr, 16), ldn = round_up(n, 16);
  const size_t r_count = rstruct == CAPITAL_UPPERTRI_PACKED ? (size_t)n * (n + 1) / 2 : (size_t)n * n;
  const double *dA, *dQ, *dRin;
  CAP_TRY(cap_stage_in(ctx, A_local, (size_t)lr * n, "A_in", &dA));
  CAP_TRY(cap_stage_in(ctx, Q_local, (size_t)lr * n, "Q_in", &dQ));
  CAP_TRY(cap_stage_in(ctx, R_local, r_count, "R_in", &dRin));
  double *Q, *Qt, *Et, *R, *G;
  CAP_TRY(ctx->workspace("qrQ", (size_t)ldq * n * 8, (void**)&Q));
  CAP_TRY(ctx->workspace("qrQt", (size_t)ldn * lr * 8, (void**)&Qt));
  CAP_TRY(ctx->workspace("qrQt2", (size_t)ldn * lr * 8, (void**)&Et));
  CAP_TRY(ctx->workspace("qrR1", (size_t)ldn * n * 8, (void**)&R));
  CAP_TRY(ctx->workspace("qrG", (size_t)ldn * n * 8, (void**)&G));
  CAP_TRY(copy_block(ctx, st, lr, n, dQ, lr, Q, ldq));
  if (rstruct == CAPITAL_UPPERTRI_PACKED) CAP_TRY(unpack_upper(ctx, st, n, dRin, R, ldn));
  else CAP_TRY(triu_copy(ctx, st, n, dRin, n, R, ldn, 0));
  CAP_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, 3 * sizeof(double), st));
  // residual (validate.hpp:37-52): ||QR - A||_F / ||A||_F, via (QR)^T - A^T = R^T Q^T - A^T
  CAP_TRY(transpose_block(ctx, st, lr, n, Q, ldq, Qt, ldn, 1.0));
  CAP_TRY(transpose_block(ctx, st, lr, n, dA, lr, Et, ldn, 1.0));
  CAP_TRY(sumsq_block(ctx, st, n, lr, Et, ldn, 0, 0, 0, 1, ctx->d_scalars + 1));
  CAP_TRY(gemm_tn(ctx, st, n, lr, n, 1.0, R, ldn, Qt, ldn, -1.0, Et, ldn, CAPITAL_GEMM_A_UPPER));
  CAP_TRY(sumsq_block(ctx, st, n, lr, Et, ldn, 0, 0, 0, 1, ctx->d_scalars));
  // orthogonality (validate.hpp:7-35): ||Q^T Q - I||_F / sqrt(n^2)
  CAP_CUDA(cudaMemsetAsync(G, 0, (size_t)ldn * n * 8, st));
  CAP_TRY(gemm_tn_splitk(ctx, st, n, n, lr, 1.0, Q, ldq, Q, ldq, G, ldn, 0));
  if (g.size > 1) CAP_NCCL(nccl().AllReduce(G, G, (size_t)ldn * n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm_world, st));
  CAP_TRY(sub_identity_local(ctx, st, n, G, ldn));
  CAP_TRY(sumsq_block(ctx, st, n, n, G, ldn, 0, 0, 0, 1, ctx->d_scalars + 2));
  CAP_TRY(allreduce_scalars(ctx, ctx->d_scalars, 2));  // numerator/denominator of the residual are row-partitioned sums
  double h[3];
  CAP_CUDA(cudaMemcpyAsync(h, ctx->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CAP_CUDA(cudaStreamSynchronize(st));
  *residual = sqrt(h[0]) / sqrt(h[1]);
  *orthogonality = sqrt(h[2]) / sqrt((double)n * (double)n);
  return CAPITAL_OK;
}
