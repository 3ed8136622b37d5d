// Krylov BLAS-1: bandwidth-bound streaming kernels over flat block-tensor buffers (SURVEY.md 8a row a7,
// kernel family K3).  128-bit vectorised coalesced loads, warp-shuffle + shared-memory block reduction,
// deterministic two-pass grid reduction (no atomics), reduction results stay on the device.
#include "internal.h"

namespace mpsk {

constexpr int KV = 8;              // basis vectors handled per pass
constexpr int RED_THREADS = 256;

struct VecPtrs { const double2* v[KV]; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[(block, j)] = sum over this block's elements of conj(V_j) * w   (complex) or V_j * w (real, packed 2 per double2)
template <bool CPLX, int NV>
__global__ void __launch_bounds__(RED_THREADS) multidot_kernel(VecPtrs V, const double2* __restrict__ w, int64_t n2,
                                                               double* __restrict__ partial) {
  double re[NV], im[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) { re[j] = 0.0; im[j] = 0.0; }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 x = w[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double2 v = V.v[j][i];
      if (CPLX) { re[j] += v.x * x.x + v.y * x.y; im[j] += v.x * x.y - v.y * x.x; }
      else { re[j] += v.x * x.x + v.y * x.y; }
    }
  }
  __shared__ double sm[RED_THREADS / 32][2 * NV];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double a = warp_sum(re[j]);
    const double b = CPLX ? warp_sum(im[j]) : 0.0;
    if (lane == 0) { sm[warp][2 * j] = a; sm[warp][2 * j + 1] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < RED_THREADS / 32; ++wv) s += sm[wv][threadIdx.x];
    partial[(int64_t)blockIdx.x * 2 * NV + threadIdx.x] = s;
  }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nblocks, int width, int ncols, double* __restrict__ out) {
  // one warp per output column; deterministic order
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= ncols) return;
  double s = 0.0;
  for (int b = threadIdx.x & 31; b < nblocks; b += 32) s += partial[(int64_t)b * width + col];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[col] = s;
}

template <bool CPLX, int NV>
__global__ void __launch_bounds__(256) multiaxpy_kernel(VecPtrs V, const double* __restrict__ coef, double sign,
                                                        double2* __restrict__ w, int64_t n2) {
  double cr[NV], ci[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) { cr[j] = sign * coef[2 * j]; ci[j] = sign * coef[2 * j + 1]; }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 x = w[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double2 v = V.v[j][i];
      if (CPLX) { x.x += cr[j] * v.x - ci[j] * v.y; x.y += cr[j] * v.y + ci[j] * v.x; }
      else { x.x += cr[j] * v.x; x.y += cr[j] * v.y; }
    }
    w[i] = x;
  }
}

__global__ void __launch_bounds__(256) scale_rsqrt_kernel(double2* __restrict__ x, int64_t n2, const double* __restrict__ nrm2) {
  const double s = rsqrt(*nrm2);
  const double s2 = s * (1.5 - 0.5 * (*nrm2) * s * s);   // one Newton step: rsqrt() is not correctly rounded
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 v = x[i]; v.x *= s2; v.y *= s2; x[i] = v;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) scale_kernel(double2* __restrict__ x, int64_t n2, double ar, double ai) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 v = x[i];
    if (CPLX) x[i] = make_double2(ar * v.x - ai * v.y, ar * v.y + ai * v.x);
    else x[i] = make_double2(ar * v.x, ar * v.y);
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) axpy_kernel(double ar, double ai, const double2* __restrict__ x, double2* __restrict__ y, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = x[i]; double2 o = y[i];
    if (CPLX) { o.x += ar * v.x - ai * v.y; o.y += ar * v.y + ai * v.x; }
    else { o.x += ar * v.x; o.y += ar * v.y; }
    y[i] = o;
  }
}

__global__ void add_small_kernel(int n, const double* a, double* b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] += a[i];
}

// number of double2 items of a flat buffer of nelem elements
static inline int64_t n2_of(bool cplx, int64_t nelem) { return cplx ? nelem : nelem / 2; }   // nelem is a multiple of GROUP_ALIGN

static int grid_for(Ctx* ctx, int64_t n2, int threads) {
  int64_t blocks = (n2 + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)ctx->num_sms * 4));
}

static int ensure_scratch(Ctx* ctx, size_t bytes) {
  if (ctx->d_scratch_bytes >= bytes) return MPSK_OK;
  if (ctx->d_scratch) ctx->free(ctx->d_scratch);
  ctx->d_scratch = (double*)ctx->alloc(bytes);
  if (!ctx->d_scratch) return MPSK_ENOMEM;
  ctx->d_scratch_bytes = bytes;
  return MPSK_OK;
}

template <bool CPLX>
static int multidot_impl(Ctx* ctx, int nv, const void* const* V, const void* w, int64_t nelem, double* d_out) {
  const int64_t n2 = n2_of(CPLX, nelem);
  const int grid = grid_for(ctx, n2, RED_THREADS);
  MPSK_TRY(ensure_scratch(ctx, (size_t)grid * 2 * KV * sizeof(double)));
  for (int j0 = 0; j0 < nv; j0 += KV) {
    const int cnt = std::min(KV, nv - j0);
    VecPtrs P;
    for (int j = 0; j < KV; ++j) P.v[j] = (const double2*)V[j0 + std::min(j, cnt - 1)];
    int width;
    switch (cnt) {
      case 1: multidot_kernel<CPLX, 1><<<grid, RED_THREADS, 0, ctx->stream>>>(P, (const double2*)w, n2, ctx->d_scratch); width = 2; break;
      case 2: multidot_kernel<CPLX, 2><<<grid, RED_THREADS, 0, ctx->stream>>>(P, (const double2*)w, n2, ctx->d_scratch); width = 4; break;
      case 3: case 4: multidot_kernel<CPLX, 4><<<grid, RED_THREADS, 0, ctx->stream>>>(P, (const double2*)w, n2, ctx->d_scratch); width = 8; break;
      default: multidot_kernel<CPLX, 8><<<grid, RED_THREADS, 0, ctx->stream>>>(P, (const double2*)w, n2, ctx->d_scratch); width = 16; break;
    }
    // only the first 2*cnt columns are meaningful
    const int outw = 2 * cnt;
    // reduce: partial has `width` columns per block
    {
      // copy-free: reduce columns [0, outw) with stride `width`
      reduce_partials_kernel<<<(outw + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_scratch, grid, width, outw, d_out + 2 * j0);
    }
    ctx->launches += 2;
    (void)outw;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int vec_multidot(Ctx* ctx, bool cplx, int nv, const void* const* V, const void* w, int64_t nelem, double* d_out) {
  return cplx ? multidot_impl<true>(ctx, nv, V, w, nelem, d_out) : multidot_impl<false>(ctx, nv, V, w, nelem, d_out);
}

template <bool CPLX>
static int multiaxpy_impl(Ctx* ctx, int nv, const void* const* V, const double* d_coef, double sign, void* w, int64_t nelem) {
  const int64_t n2 = n2_of(CPLX, nelem);
  const int grid = grid_for(ctx, n2, 256);
  for (int j0 = 0; j0 < nv; j0 += KV) {
    const int cnt = std::min(KV, nv - j0);
    VecPtrs P;
    for (int j = 0; j < KV; ++j) P.v[j] = (const double2*)V[j0 + std::min(j, cnt - 1)];
    const double* c = d_coef + 2 * j0;
    switch (cnt) {
      case 1: multiaxpy_kernel<CPLX, 1><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 2: multiaxpy_kernel<CPLX, 2><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 3: multiaxpy_kernel<CPLX, 3><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 4: multiaxpy_kernel<CPLX, 4><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 5: multiaxpy_kernel<CPLX, 5><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 6: multiaxpy_kernel<CPLX, 6><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      case 7: multiaxpy_kernel<CPLX, 7><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
      default: multiaxpy_kernel<CPLX, 8><<<grid, 256, 0, ctx->stream>>>(P, c, sign, (double2*)w, n2); break;
    }
    ctx->launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int vec_multiaxpy(Ctx* ctx, bool cplx, int nv, const void* const* V, const double* d_coef, double sign, void* w, int64_t nelem) {
  return cplx ? multiaxpy_impl<true>(ctx, nv, V, d_coef, sign, w, nelem) : multiaxpy_impl<false>(ctx, nv, V, d_coef, sign, w, nelem);
}

int vec_norm2(Ctx* ctx, bool cplx, const void* x, int64_t nelem, double* d_out) {
  // <x|x>: reuse multidot with V = {x}; imaginary slot is written next to it, so d_out needs 2 doubles
  const void* V[1] = {x};
  return vec_multidot(ctx, cplx, 1, V, x, nelem, d_out);
}

int vec_scale_rsqrt(Ctx* ctx, bool cplx, void* x, int64_t nelem, const double* d_norm2) {
  const int64_t n2 = n2_of(cplx, nelem);
  scale_rsqrt_kernel<<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>((double2*)x, n2, d_norm2);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int vec_scale(Ctx* ctx, bool is_cplx, void* x, int64_t nelem, cplx a) {
  const int64_t n2 = n2_of(is_cplx, nelem);
  if (is_cplx) scale_kernel<true><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>((double2*)x, n2, a.real(), a.imag());
  else scale_kernel<false><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>((double2*)x, n2, a.real(), 0.0);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int vec_axpy(Ctx* ctx, bool is_cplx, cplx a, const void* x, void* y, int64_t nelem) {
  const int64_t n2 = n2_of(is_cplx, nelem);
  if (is_cplx) axpy_kernel<true><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>(a.real(), a.imag(), (const double2*)x, (double2*)y, n2);
  else axpy_kernel<false><<<grid_for(ctx, n2, 256), 256, 0, ctx->stream>>>(a.real(), 0.0, (const double2*)x, (double2*)y, n2);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int vec_add_dev(Ctx* ctx, int n, const double* a, double* b) {
  add_small_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, a, b);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

}  // namespace mpsk
