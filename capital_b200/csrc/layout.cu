// HBM-bound helpers around the GEMM kernel: window copies / transposes (what serialize<S,D>::invoke does on
// the CPU, serialize.hpp:12-150, reduced to the few that remain once operands are used in place), packed
// upper <-> rect conversion (structure.h:37-39), the reference's generators (structure.hpp:69-129) and the
// validators' Frobenius reductions (util.hpp:25-53).  All kernels are coalesced along the contiguous
// (row) index and sized in multiples of the SM count where the extent allows.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int TP = 32;

__global__ void transpose_kernel(int rows, int cols, const double* src, long long lds, double* dst,
                                 long long ldd, double scale) {
  __shared__ double tile[TP][TP + 1];
  const int r0 = blockIdx.x * TP, c0 = blockIdx.y * TP;
  for (int j = threadIdx.y; j < TP; j += blockDim.y) {
    const int r = r0 + threadIdx.x, c = c0 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(long long)c * lds + r];
  }
  __syncthreads();
  // dst is cols x rows: dst(c, r) = src(r, c)
  for (int j = threadIdx.y; j < TP; j += blockDim.y) {
    const int c = c0 + threadIdx.x, r = r0 + j;
    if (r < rows && c < cols) dst[(long long)r * ldd + c] = scale * tile[threadIdx.x][j];
  }
}

__global__ void copy_kernel(long long rows, long long cols, const double* src, long long lds, double* dst,
                            long long ldd) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / rows, r = i - c * rows;
    dst[c * ldd + r] = src[c * lds + r];
  }
}

__global__ void zero_kernel(long long rows, long long cols, double* dst, long long ldd) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / rows, r = i - c * rows;
    dst[c * ldd + r] = 0.0;
  }
}

// one block per column chunk: column i of the packed triangle is contiguous (i+1 entries at i(i+1)/2)
__global__ void pack_upper_kernel(long long c0, long long n, const double* src, long long lds, double* packed,
                                  int zero_diag) {
  for (long long i = c0 + blockIdx.x; i < n; i += gridDim.x) {
    const double* s = src + i * lds;
    double* d = packed + i * (i + 1) / 2;
    for (long long j = threadIdx.x; j <= i; j += blockDim.x) d[j] = (zero_diag && j == i) ? 0.0 : s[j];
  }
}
// rows [r0, min(r1, col + 1)) of columns [c0, c1) of a rect matrix -> their slots of the packed upper triangle.  `packed` may be
// the device alias of a pinned host array: each column is one contiguous run of coalesced 8-byte stores, a handful of CTAs
// keeps the PCIe link busy.
__global__ void emit_block_packed_kernel(const double* src, long long lds, double* packed, long long r0, long long r1,
                                         long long c0, long long c1) {
  for (long long i = c0 + blockIdx.x; i < c1; i += gridDim.x) {
    const double* s = src + i * lds;
    double* d = packed + i * (i + 1) / 2;
    const long long re = r1 < i + 1 ? r1 : i + 1;
    for (long long j = r0 + threadIdx.x; j < re; j += blockDim.x) d[j] = s[j];
  }
}
__global__ void unpack_upper_kernel(long long n, const double* packed, double* dst, long long ldd) {
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
    const double* s = packed + i * (i + 1) / 2;
    double* d = dst + i * ldd;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) d[j] = j <= i ? s[j] : 0.0;
  }
}
__global__ void triu_copy_kernel(long long n, const double* src, long long lds, double* dst, long long ldd,
                                 int zero_diag) {
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
    const double* s = src + i * lds;
    double* d = dst + i * ldd;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) d[j] = (j < i || (j == i && !zero_diag)) ? s[j] : 0.0;
  }
}

// drand48: X0 = seed<<16 | 0x330E ; X1 = (a X0 + c) mod 2^48 ; value = X1 / 2^48  (structure.hpp:80-85 re-seeds per element)
__device__ __forceinline__ double drand48_first(unsigned long long seed) {
  const unsigned long long a = 0x5DEECE66DULL, c = 0xBULL, m48 = (1ULL << 48) - 1;
  unsigned long long x = ((seed & 0xFFFFFFFFULL) << 16) | 0x330EULL;
  x = (a * x + c) & m48;  // 64-bit wraparound keeps the low 48 bits exact
  return (double)x * (1.0 / 281474976710656.0);
}

__global__ void gen_symmetric_kernel(double* A, long long ld, long long lrows, long long lcols, long long n, int x, int y,
                                     int d, int diag_dom) {
  const long long total = lrows * lcols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / lrows, j = idx - i * lrows;  // local col i, local row j
    const long long gx = x + i * d, gy = y + j * d;
    double v = 0.0;
    if (gx < n && gy < n) {
      const unsigned long long hi = gx > gy ? gx : gy, lo = gx > gy ? gy : gx;
      v = drand48_first(hi + (unsigned long long)n * lo);
      if (diag_dom && gx == gy && i == j) v += (double)n;
    }
    A[i * ld + j] = v;
  }
}

// draw number t (0-based) of the stream after srand48(key): X_{t+1} = a^{t+1} X0 + c (a^{t+1}-1)/(a-1)
__device__ __forceinline__ void lcg_pow(unsigned long long t, unsigned long long& A, unsigned long long& C) {
  const unsigned long long m48 = (1ULL << 48) - 1;
  unsigned long long ca = 0x5DEECE66DULL, cc = 0xBULL;  // current step (a, c) for 2^bit
  A = 1; C = 0;
  while (t) {
    if (t & 1) { A = (A * ca) & m48; C = (C * ca + cc) & m48; }
    cc = (cc * ca + cc) & m48;
    ca = (ca * ca) & m48;
    t >>= 1;
  }
}
__global__ void gen_random_kernel(double* Aout, long long ld, long long lrows, long long lcols, long long pad_rows,
                                  long long pad_cols, long long key) {
  const unsigned long long m48 = (1ULL << 48) - 1;
  const unsigned long long x0 = (((unsigned long long)key & 0xFFFFFFFFULL) << 16) | 0x330EULL;
  const long long total = lrows * lcols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / lrows, j = idx - i * lrows;
    double v = 0.0;
    if (i < pad_cols && j < pad_rows) {
      unsigned long long A, C;
      lcg_pow((unsigned long long)(i * pad_rows + j) + 1ULL, A, C);
      const unsigned long long xv = (A * x0 + C) & m48;
      v = (double)xv * (1.0 / 281474976710656.0);
    }
    Aout[i * ld + j] = v;
  }
}

// upper_mode: 0 = all entries, 1 = only entries whose GLOBAL position satisfies row <= col
__global__ void sumsq_kernel(long long rows, long long cols, const double* a, long long ld, int upper_mode, int x, int y,
                             int d, double* out) {
  double s = 0.0;
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / rows, j = idx - i * rows;
    if (upper_mode && (y + j * d) > (x + i * d)) continue;
    const double v = a[i * ld + j];
    s += v * v;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void sub_identity_kernel(long long n, double* a, long long ld) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i * ld + i] -= 1.0;
}

// zero the band |row - col| <= hw of an n x n matrix
__global__ void zero_band_kernel(long long n, long long hw, double* a, long long ld) {
  const long long w = 2 * hw + 1, total = n * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / w, r = c - hw + (idx - c * w);
    if (r >= 0 && r < n) a[c * ld + r] = 0.0;
  }
}

inline int grid_for(const capital_ctx* ctx, long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)ctx->num_sms * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

#define LAUNCH_CHECK()                      \
  do {                                      \
    ctx->counters.kernel_launches++;        \
    CAP_CUDA(cudaGetLastError());           \
  } while (0)

capital_status_t transpose_block(capital_ctx* ctx, cudaStream_t st, int64_t rows, int64_t cols, const double* src, int64_t lds,
                                 double* dst, int64_t ldd, double scale) {
  if (rows <= 0 || cols <= 0) return CAPITAL_OK;
  dim3 grid((unsigned)ceil_div(rows, TP), (unsigned)ceil_div(cols, TP)), block(TP, 8);
  const int tli = ctx->tl_begin(st, 8, 1, (double)rows, (double)cols);
  transpose_kernel<<<grid, block, 0, st>>>((int)rows, (int)cols, src, lds, dst, ldd, scale);
  ctx->tl_end(st, tli);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t copy_block(capital_ctx* ctx, cudaStream_t st, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
                            int64_t ldd) {
  if (rows <= 0 || cols <= 0) return CAPITAL_OK;
  if (lds == rows && ldd == rows) {
    const size_t total = (size_t)rows * cols * 8, piece = (size_t)1 << 30;  // pieces of at most 1 GiB (see dist.cu: dma2d)
    for (size_t off = 0; off < total; off += piece)
      CAP_CUDA(cudaMemcpyAsync((char*)dst + off, (const char*)src + off, std::min(piece, total - off), cudaMemcpyDeviceToDevice, st));
    return CAPITAL_OK;
  }
  const int tli = ctx->tl_begin(st, 8, 2, (double)rows, (double)cols);
  copy_kernel<<<grid_for(ctx, rows * cols, 256), 256, 0, st>>>(rows, cols, src, lds, dst, ldd);
  ctx->tl_end(st, tli);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t zero_block(capital_ctx* ctx, cudaStream_t st, int64_t rows, int64_t cols, double* dst, int64_t ldd) {
  if (rows <= 0 || cols <= 0) return CAPITAL_OK;
  if (ldd == rows) {
    CAP_CUDA(cudaMemsetAsync(dst, 0, (size_t)rows * cols * 8, st));
    return CAPITAL_OK;
  }
  zero_kernel<<<grid_for(ctx, rows * cols, 256), 256, 0, st>>>(rows, cols, dst, ldd);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t pack_upper(capital_ctx* ctx, cudaStream_t st, int64_t n, const double* src, int64_t lds, double* packed, int zero_diag,
                            int64_t col_begin, int64_t col_end) {
  if (col_end < 0) col_end = n;
  const int64_t cols = col_end - col_begin;
  if (cols <= 0) return CAPITAL_OK;
  const int tli = ctx->tl_begin(st, 8, 3, (double)col_begin, (double)col_end);
  pack_upper_kernel<<<(int)(cols < ctx->num_sms * 8 ? cols : ctx->num_sms * 8), 256, 0, st>>>(col_begin, col_end, src, lds, packed, zero_diag);
  ctx->tl_end(st, tli);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t emit_block_packed(capital_ctx* ctx, cudaStream_t st, const double* src, int64_t lds, double* packed, int64_t r0, int64_t r1,
                                   int64_t c0, int64_t c1, int ctas) {
  if (c1 <= c0 || r1 <= r0) return CAPITAL_OK;
  const int64_t cols = c1 - c0;
  if (ctas < 1) ctas = 1;
  emit_block_packed_kernel<<<(int)(cols < ctas ? cols : ctas), 512, 0, st>>>(src, lds, packed, r0, r1, c0, c1);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t unpack_upper(capital_ctx* ctx, cudaStream_t st, int64_t n, const double* packed, double* dst, int64_t ldd) {
  if (n <= 0) return CAPITAL_OK;
  unpack_upper_kernel<<<(int)(n < ctx->num_sms * 8 ? n : ctx->num_sms * 8), 256, 0, st>>>(n, packed, dst, ldd);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t triu_copy(capital_ctx* ctx, cudaStream_t st, int64_t n, const double* src, int64_t lds, double* dst, int64_t ldd,
                           int zero_diag) {
  if (n <= 0) return CAPITAL_OK;
  triu_copy_kernel<<<(int)(n < ctx->num_sms * 8 ? n : ctx->num_sms * 8), 256, 0, st>>>(n, src, lds, dst, ldd, zero_diag);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t gen_symmetric(capital_ctx* ctx, cudaStream_t st, double* A, int64_t ld, int64_t lrows, int64_t lcols, int64_t n_global,
                               int x, int y, int d, int diag_dom) {
  gen_symmetric_kernel<<<grid_for(ctx, lrows * lcols, 256), 256, 0, st>>>(A, ld, lrows, lcols, n_global, x, y, d, diag_dom);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t gen_random(capital_ctx* ctx, cudaStream_t st, double* A, int64_t ld, int64_t lrows, int64_t lcols, int64_t pad_rows,
                            int64_t pad_cols, int64_t key) {
  gen_random_kernel<<<grid_for(ctx, lrows * lcols, 256), 256, 0, st>>>(A, ld, lrows, lcols, pad_rows, pad_cols, key);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t sumsq_block(capital_ctx* ctx, cudaStream_t st, int64_t rows, int64_t cols, const double* a, int64_t ld, int upper_mode,
                             int x, int y, int d, double* out) {
  if (rows <= 0 || cols <= 0) return CAPITAL_OK;
  sumsq_kernel<<<grid_for(ctx, rows * cols, 256), 256, 0, st>>>(rows, cols, a, ld, upper_mode, x, y, d, out);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
capital_status_t sub_identity_local(capital_ctx* ctx, cudaStream_t st, int64_t n, double* a, int64_t ld) {
  sub_identity_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(n, a, ld);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}

// The triangular products read whole diagonal GEMM tiles (<= 128 wide) of Rinv / Rinv^T, including entries on the other side of
// the diagonal that no kernel writes; everything farther than one tile from the diagonal is either written before it is read
// or never read.  Zeroing the band |row - col| <= 256 therefore replaces a memset of the whole n x n buffer.
capital_status_t zero_band(capital_ctx* ctx, cudaStream_t st, int64_t n, double* a, int64_t ld) {
  if (n <= 0) return CAPITAL_OK;
  const int64_t hw = 256;
  if (n <= 4 * hw) {
    CAP_CUDA(cudaMemsetAsync(a, 0, (size_t)ld * n * 8, st));
    return CAPITAL_OK;
  }
  zero_band_kernel<<<grid_for(ctx, n * (2 * hw + 1), 256), 256, 0, st>>>(n, hw, a, ld);
  LAUNCH_CHECK();
  return CAPITAL_OK;
}
