// Path B: RolX NMF, multiplicative updates for the Frobenius loss, on sm_100a.
//
// Replaces sklearn.decomposition._nmf._fit_multiplicative_update (beta_loss = 2, no
// regularisation; sklearn/decomposition/_nmf.py:726-888) as reached from
// graphrole/roles/factor.py:19-25.  One iteration is
//     W <- W * (X H^T) / (W (H H^T))          (_nmf.py:535-549, :615-624)
//     H <- H * (W^T X) / ((W^T W) H)          (_nmf.py:633-635, :701-721; uses the new W)
// with exactly-zero denominators replaced by float32 eps (_nmf.py:32).
//
// This file holds the host loop, the small r x r / r x f kernels shared by both compute paths,
// and the fp32 FFMA path (use_tf32 == 0, and the shapes the tcgen05 kernel does not take):
//   nmf_update_w_kernel   X H^T per 64-row block with the W update fused in the epilogue
//   nmf_wt_x_kernel       W^T X (and W^T W) as per-split partials, reduced in fixed order
//   nmf_update_h_kernel   partial reduction + H update;   nmf_hht_kernel  H H^T
//   nmf_error_kernel      ||X - W H||_F^2 from the dense residual (_nmf.py:122), fp64 sums
// The tcgen05 / TMA fused single-pass kernel lives in nmf_mu_tc.cu.

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "nmf_handle.cuh"

using namespace gr;

namespace {

constexpr float kEps = 1.1920928955078125e-07f;  // np.finfo(np.float32).eps, _nmf.py:32
constexpr int kThreads = 256;
constexpr int kBM = 64;  // rows per CTA tile in nmf_update_w_kernel
constexpr int kBK = 32;  // K (feature) chunk

// ---- H H^T (r x r): one CTA per row i, 8 column slices per output, fixed-order combine ----------
__global__ void __launch_bounds__(256)
nmf_hht_kernel(const float* __restrict__ H, int r, int f, float* __restrict__ HHt) {
    __shared__ float part[8][33];
    const int i = blockIdx.x, j = threadIdx.x % 32, slice = threadIdx.x / 32;
    float acc = 0.f;
    if (j < r)
        for (int c = slice; c < f; c += 8)
            acc = fmaf(H[(int64_t)i * f + c], H[(int64_t)j * f + c], acc);
    part[slice][j] = acc;
    __syncthreads();
    if (slice == 0 && j < r) {
        float t = 0.f;
        for (int s = 0; s < 8; ++s) t += part[s][j];
        HHt[i * r + j] = t;
    }
}

// ---- W update: XHt tile on FFMA, epilogue W *= XHt / (W HHt) ----------------------------------
template <int RP>
__global__ void __launch_bounds__(kThreads)
nmf_update_w_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int f, int r,
                    const float* __restrict__ H, const float* __restrict__ HHt,
                    float* __restrict__ W) {
    constexpr int RQ = RP / 4;  // roles per thread
    __shared__ float Xs[kBM][kBK + 1];
    __shared__ float Hs[RP][kBK + 1];
    __shared__ float Ws[kBM][RP + 1];
    __shared__ float HHts[RP][RP + 1];
    const int tid = threadIdx.x;
    const int row = tid / 4, rg = tid % 4;
    const int64_t row0 = (int64_t)blockIdx.x * kBM;

    for (int i = tid; i < RP * RP; i += kThreads) {
        const int a = i / RP, b = i % RP;
        HHts[a][b] = (a < r && b < r) ? HHt[a * r + b] : 0.f;
    }
    for (int i = tid; i < kBM * RP; i += kThreads) {
        const int a = i / RP, b = i % RP;
        Ws[a][b] = (row0 + a < n && b < r) ? W[(row0 + a) * r + b] : 0.f;
    }

    float acc[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) acc[j] = 0.f;

    for (int k0 = 0; k0 < f; k0 += kBK) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kBM * kBK / kThreads; ++i) {
            const int rr = i * (kThreads / kBK) + tid / kBK, cc = tid % kBK;
            Xs[rr][cc] = (row0 + rr < n && k0 + cc < f) ? __ldg(X + (row0 + rr) * ldx + k0 + cc)
                                                         : 0.f;
        }
        for (int i = tid; i < RP * kBK; i += kThreads) {
            const int a = i / kBK, cc = i % kBK;
            Hs[a][cc] = (a < r && k0 + cc < f) ? __ldg(H + (int64_t)a * f + k0 + cc) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kBK; ++kk) {
            const float x = Xs[row][kk];
#pragma unroll
            for (int j = 0; j < RQ; ++j) acc[j] = fmaf(x, Hs[rg * RQ + j][kk], acc[j]);
        }
    }
    if (row0 + row >= n) return;
#pragma unroll
    for (int j = 0; j < RQ; ++j) {
        const int role = rg * RQ + j;
        if (role >= r) continue;
        float den = 0.f;
        for (int l = 0; l < r; ++l) den = fmaf(Ws[row][l], HHts[l][role], den);
        if (den == 0.f) den = kEps;
        W[(row0 + row) * r + role] = Ws[row][role] * (acc[j] / den);
    }
}

// ---- partial W^T M for a row-major M [n, fc] (M = X gives W^T X, M = W gives W^T W) -----------
// grid: (column slabs of kThreads, row splits).  out[(split * RP + l) * fc + col]
template <int RP>
__global__ void __launch_bounds__(kThreads)
nmf_wt_x_kernel(const float* __restrict__ M, int64_t ldm, int64_t n, int fc,
                const float* __restrict__ W, int r, int64_t rows_per_split,
                float* __restrict__ out) {
    constexpr int RB = 32;  // rows staged per step
    __shared__ float Ws[RB][RP];
    const int col = blockIdx.x * kThreads + threadIdx.x;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(n, r_lo + rows_per_split);
    float acc[RP];
#pragma unroll
    for (int l = 0; l < RP; ++l) acc[l] = 0.f;

    for (int64_t b = r_lo; b < r_hi; b += RB) {
        __syncthreads();
        for (int i = threadIdx.x; i < RB * RP; i += kThreads) {
            const int a = i / RP, l = i % RP;
            Ws[a][l] = (b + a < r_hi && l < r) ? __ldg(W + (b + a) * r + l) : 0.f;
        }
        __syncthreads();
        if (col < fc) {
            const int cnt = (int)min((int64_t)RB, r_hi - b);
            float x[8];
            for (int a0 = 0; a0 < cnt; a0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    x[u] = a0 + u < cnt ? __ldg(M + (b + a0 + u) * ldm + col) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int l = 0; l < RP; ++l) acc[l] = fmaf(Ws[a0 + u][l], x[u], acc[l]);
            }
        }
    }
    if (col < fc)
#pragma unroll
        for (int l = 0; l < RP; ++l)
            if (l < r) out[((int64_t)blockIdx.y * RP + l) * fc + col] = acc[l];
}

// ---- fixed-order reduction of the W^T W partials (r x r) --------------------------------------
__global__ void __launch_bounds__(1024)
nmf_reduce_wtw_kernel(const float* __restrict__ part, int splits, int rp, int r,
                      float* __restrict__ WtW) {
    const int i = threadIdx.x / 32, j = threadIdx.x % 32;
    if (i >= r || j >= r) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[((int64_t)s * rp + i) * r + j];
    WtW[i * r + j] = acc;
}

// ---- fixed-order reduction of the W^T X partials (r x f): grid (column slabs, r) ---------------
__global__ void __launch_bounds__(kThreads)
nmf_reduce_wtx_kernel(const float* __restrict__ part, int splits, int rp, int f,
                      float* __restrict__ WtX) {
    const int col = blockIdx.x * kThreads + threadIdx.x;
    const int l = blockIdx.y;
    if (col >= f) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[((int64_t)s * rp + l) * f + col];
    WtX[(int64_t)l * f + col] = acc;
}

// ---- H update: reduce W^T X partials, H_next = H * WtX / (WtW H) --------------------------------
// One thread per (role, column): grid (column slabs, r).  Reads the old H, writes H_next (the
// caller copies it back), so no thread sees a half-updated column.
__global__ void __launch_bounds__(kThreads)
nmf_update_h_kernel(const float* __restrict__ part, int splits, int rp, int r, int f,
                    const float* __restrict__ WtW, const float* __restrict__ H,
                    float* __restrict__ H_next) {
    const int col = blockIdx.x * kThreads + threadIdx.x;
    const int l = blockIdx.y;
    if (col >= f) return;
    float num = 0.f;
    for (int s = 0; s < splits; ++s) num += part[((int64_t)s * rp + l) * f + col];  // fixed order
    float den = 0.f;
    for (int m = 0; m < r; ++m) den = fmaf(__ldg(WtW + l * r + m), H[(int64_t)m * f + col], den);
    if (den == 0.f) den = kEps;
    H_next[(int64_t)l * f + col] = H[(int64_t)l * f + col] * (num / den);
}

// ---- ||X - W H||_F^2 partials: thread per column, fp64 accumulation ----------------------------
template <int RP>
__global__ void __launch_bounds__(kThreads)
nmf_error_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int f,
                 const float* __restrict__ W, const float* __restrict__ H, int r,
                 int64_t rows_per_split, double* __restrict__ out) {
    constexpr int RB = 32;
    __shared__ __align__(16) float Ws[RB][RP];
    __shared__ double red[kThreads / 32];
    const int col = blockIdx.x * kThreads + threadIdx.x;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(n, r_lo + rows_per_split);
    float h[RP];
#pragma unroll
    for (int l = 0; l < RP; ++l) h[l] = (l < r && col < f) ? __ldg(H + (int64_t)l * f + col) : 0.f;
    double total = 0.0;
    for (int64_t b = r_lo; b < r_hi; b += RB) {
        __syncthreads();
        for (int i = threadIdx.x; i < RB * RP; i += kThreads) {
            const int a = i / RP, l = i % RP;
            Ws[a][l] = (b + a < r_hi && l < r) ? __ldg(W + (b + a) * r + l) : 0.f;
        }
        __syncthreads();
        if (col < f) {
            // sixteen rows at a time: their X loads are independent (16 requests in flight per
            // thread -- with one, the pass ran at 1.9 TB/s on C5 -- ) and W comes out of shared
            // memory as float4 broadcasts: one LDS.128 per four FMAs.  Rows past r_hi have W = 0
            // and X read as 0, so they add nothing; the fp32 partial keeps the row order.
            constexpr int UR = 16;
            float part = 0.f;
            for (int a = 0; a < RB; a += UR) {
                float x[UR], wh[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    x[u] = b + a + u < r_hi ? __ldg(X + (b + a + u) * ldx + col) : 0.f;
                    wh[u] = 0.f;
                }
#pragma unroll
                for (int l4 = 0; l4 < RP / 4; ++l4)
#pragma unroll
                    for (int u = 0; u < UR; ++u) {
                        const float4 w = *reinterpret_cast<const float4*>(&Ws[a + u][4 * l4]);
                        wh[u] = fmaf(w.x, h[4 * l4], wh[u]);
                        wh[u] = fmaf(w.y, h[4 * l4 + 1], wh[u]);
                        wh[u] = fmaf(w.z, h[4 * l4 + 2], wh[u]);
                        wh[u] = fmaf(w.w, h[4 * l4 + 3], wh[u]);
                    }
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const float diff = x[u] - wh[u];
                    part = fmaf(diff, diff, part);
                }
            }
            total += (double)part;
        }
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += red[w];
        out[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// ---- ||X - W H||_F^2, register-tiled: CTA = 128 columns, 128-row steps, 8 x 8 outputs per thread --
// The thread-per-column kernel above spends one shared-memory load per four FMAs and keeps one
// row in flight per step: on C5 it ran at 10 ms (r = 4) - 22 ms (r = 32) per pass; with 256
// threads per CTA (one CTA per SM at 233 registers) this kernel took 8.2 - 13.6 ms.  Here every
// thread holds an 8 x 8 block of W.H in registers (10 shared-memory loads per 64 FMAs), its 16
// float4 loads of X are issued before the FMAs, and the pass is bound by the X stream (small r)
// or by FFMA issue (r = 32).  Needs f % 4 == 0 and 16-byte aligned rows; the kernel above stays
// as the general form.  Rows past r_hi and columns past f contribute exactly 0.
constexpr int kErrTile = 128;       // columns per CTA
constexpr int kErrRows = 64;        // rows per step: 128 threads x (8 x 8), two CTAs per SM --
                                    // one computes while the other waits for its X loads
constexpr int kErrThreads = 128;
constexpr int kMaxErrParts = 8192;
template <int RP>
__global__ void __launch_bounds__(kErrThreads, 2)
nmf_error_tile_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int f,
                      const float* __restrict__ W, const float* __restrict__ H, int r,
                      int64_t rows_per_split, double* __restrict__ out) {
    __shared__ __align__(16) float Hs[RP][kErrTile];
    __shared__ float Ws[kErrRows][RP + 1];   // +1: the two row groups of a warp hit different banks
    __shared__ double red[kErrThreads / 32];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int col0 = (int)blockIdx.x * kErrTile, c = col0 + tx * 8;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = min(n, r_lo + rows_per_split);
    for (int i = threadIdx.x; i < RP * kErrTile; i += kErrThreads) {
        const int k = i / kErrTile, cc = i % kErrTile;
        Hs[k][cc] = (k < r && col0 + cc < f) ? __ldg(H + (int64_t)k * f + col0 + cc) : 0.f;
    }
    double total = 0.0;
    for (int64_t b = r_lo; b < r_hi; b += kErrRows) {
        __syncthreads();
        for (int i = threadIdx.x; i < kErrRows * RP; i += kErrThreads) {
            const int row = i / RP, k = i % RP;
            Ws[row][k] = (b + row < r_hi && k < r) ? __ldg(W + (b + row) * r + k) : 0.f;
        }
        __syncthreads();
        float x[8][8], acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = b + ty * 8 + i;
            const bool row_ok = row < r_hi;
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_ok && c + 4 * v < f)
                    t = __ldcs(reinterpret_cast<const float4*>(X + row * ldx + c + 4 * v));
                x[i][4 * v] = t.x; x[i][4 * v + 1] = t.y; x[i][4 * v + 2] = t.z; x[i][4 * v + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        }
#pragma unroll 4
        for (int k = 0; k < RP; ++k) {
            const float4 h0 = *reinterpret_cast<const float4*>(&Hs[k][tx * 8]);
            const float4 h1 = *reinterpret_cast<const float4*>(&Hs[k][tx * 8 + 4]);
            const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float w = Ws[ty * 8 + i][k];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w, hv[j], acc[i][j]);
            }
        }
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float diff = x[i][j] - acc[i][j];
                part = fmaf(diff, diff, part);
            }
        total += (double)part;
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kErrThreads / 32; ++w) t += red[w];
        out[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// ---- rank padding: W [n, r] <-> [n, r4], zero columns from r on -----------------------------------
__global__ void nmf_pad_w_kernel(const float* __restrict__ W, int64_t n, int r, int r4,
                                 float* __restrict__ Wp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * r4) return;
    const int64_t row = i / r4;
    const int c = (int)(i % r4);
    Wp[i] = c < r ? W[row * r + c] : 0.f;
}
__global__ void nmf_unpad_w_kernel(const float* __restrict__ Wp, int64_t n, int r, int r4,
                                   float* __restrict__ W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * r) return;
    W[i] = Wp[(i / r) * r4 + i % r];
}

template <typename F>
int dispatch_rp(int rp, F&& fn) {
    switch (rp) {
        case 8: return fn(std::integral_constant<int, 8>{});
        case 16: return fn(std::integral_constant<int, 16>{});
        default: return fn(std::integral_constant<int, 32>{});
    }
}

#define GR_LAUNCH_CHECK(name)                                                              \
    do {                                                                                   \
        count_launch();                                                                    \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return fail(GR_ERR_CUDA, "%s launch failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

}  // namespace

int gr::nmf_hht(gr_nmf* h, const float* H, cudaStream_t st) {
    nmf_hht_kernel<<<h->r, 256, 0, st>>>(H, h->r, h->f, h->d_hht);
    GR_LAUNCH_CHECK("nmf_hht_kernel");
    return GR_OK;
}

// ---- FFMA path: one multiplicative-update iteration --------------------------------------------
int gr::nmf_iteration_fma(gr_nmf* h, const float* X, int64_t ldx, float* W, float* H,
                          cudaStream_t st) {
    const int r = h->r, f = h->f, rp = h->rp;
    const int64_t n = h->n;
    if (int rc = nmf_hht(h, H, st)) return rc;
    if (int rc = dispatch_rp(rp, [&](auto RPc) {
            constexpr int RP = decltype(RPc)::value;
            nmf_update_w_kernel<RP><<<(unsigned)ceil_div<int64_t>(n, kBM), kThreads, 0, st>>>(
                X, ldx, n, f, r, H, h->d_hht, W);
            GR_LAUNCH_CHECK("nmf_update_w_kernel");
            dim3 gx((unsigned)ceil_div(f, kThreads), (unsigned)h->splits);
            nmf_wt_x_kernel<RP><<<gx, kThreads, 0, st>>>(X, ldx, n, f, W, r, h->rows_per_split,
                                                        h->d_part_wtx);
            GR_LAUNCH_CHECK("nmf_wt_x_kernel(X)");
            dim3 gw(1, (unsigned)h->splits);
            nmf_wt_x_kernel<RP><<<gw, kThreads, 0, st>>>(W, r, n, r, W, r, h->rows_per_split,
                                                        h->d_part_wtw);
            GR_LAUNCH_CHECK("nmf_wt_x_kernel(W)");
            return (int)GR_OK;
        }))
        return rc;
    return nmf_finish_iteration(h, h->d_part_wtx, h->d_part_wtw, h->splits, h->rp, H, st);
}

// Shared tail of an iteration: reduce the partials, update H.
int gr::nmf_finish_iteration(gr_nmf* h, const float* part_wtx, const float* part_wtw, int splits,
                             int rp, float* H, cudaStream_t st) {
    if (h->defer_finish) {      // row-sharded: the sums still have to cross the ranks
        h->pending_wtx = part_wtx;
        h->pending_wtw = part_wtw;
        h->pending_splits = splits;
        h->pending_rp = rp;
        return GR_OK;
    }
    nmf_reduce_wtw_kernel<<<1, 1024, 0, st>>>(part_wtw, splits, rp, h->r, h->d_wtw);
    GR_LAUNCH_CHECK("nmf_reduce_wtw_kernel");
    dim3 grid((unsigned)ceil_div(h->f, kThreads), (unsigned)h->r);
    nmf_update_h_kernel<<<grid, kThreads, 0, st>>>(part_wtx, splits, rp, h->r, h->f, h->d_wtw, H,
                                                 h->d_h_next);
    GR_LAUNCH_CHECK("nmf_update_h_kernel");
    GR_CUDA_TRY(cudaMemcpyAsync(H, h->d_h_next, (size_t)h->r * h->f * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    return GR_OK;
}

int gr::nmf_error(gr_nmf* h, const float* X, int64_t ldx, const float* W, const float* H,
                  double* err, cudaStream_t st) {
    int slabs = ceil_div(h->f, kThreads);
    int splits = h->splits;
    const bool tiled = h->f % 4 == 0 && ldx % 4 == 0 && aligned16(X) && !getenv("GR_NMF_ERROR_SIMPLE");
    if (tiled) {
        // 128-column blocks x row splits of whole 128-row steps: about two CTAs per SM in total
        slabs = ceil_div(h->f, kErrTile);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const int64_t steps = ceil_div<int64_t>(h->n, kErrRows);
        int64_t want = std::max<int64_t>(1, std::min<int64_t>(steps, (4 * sms + slabs - 1) / slabs));
        want = std::min<int64_t>(want, kMaxErrParts / slabs);
        const int64_t rows_per_split = ceil_div<int64_t>(steps, want) * kErrRows;
        splits = (int)ceil_div<int64_t>(h->n, rows_per_split);
        if (int rc = dispatch_rp(h->rp, [&](auto RPc) {
                constexpr int RP = decltype(RPc)::value;
                dim3 g((unsigned)slabs, (unsigned)splits);
                nmf_error_tile_kernel<RP><<<g, kErrThreads, 0, st>>>(X, ldx, h->n, h->f, W, H, h->r,
                                                             rows_per_split, h->d_err_part);
                GR_LAUNCH_CHECK("nmf_error_tile_kernel");
                return (int)GR_OK;
            }))
            return rc;
    } else if (int rc = dispatch_rp(h->rp, [&](auto RPc) {
            constexpr int RP = decltype(RPc)::value;
            dim3 g((unsigned)slabs, (unsigned)h->splits);
            nmf_error_kernel<RP><<<g, kThreads, 0, st>>>(X, ldx, h->n, h->f, W, H, h->r,
                                                        h->rows_per_split, h->d_err_part);
            GR_LAUNCH_CHECK("nmf_error_kernel");
            return (int)GR_OK;
        }))
        return rc;
    const size_t cnt = (size_t)slabs * splits;
    h->h_err_part.resize(cnt);
    GR_CUDA_TRY(cudaMemcpyAsync(h->h_err_part.data(), h->d_err_part, cnt * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    GR_CUDA_TRY(cudaStreamSynchronize(st));
    double total = 0.0;
    for (size_t i = 0; i < cnt; ++i) total += h->h_err_part[i];  // fixed order
    *err = std::sqrt(total);
    return GR_OK;
}

// ---- C-ABI --------------------------------------------------------------------------------------
extern "C" int gr_nmf_create(gr_nmf_t** out, int64_t n, int32_t f, int32_t r, int device) {
    GR_REQUIRE(out != nullptr, "gr_nmf_create: out is NULL");
    *out = nullptr;
    GR_REQUIRE(n >= 1 && f >= 1 && r >= 1, "gr_nmf_create: n, f, r must be positive");
    GR_REQUIRE(r <= 32, "gr_nmf_create: n_roles = %d not supported (r <= 32)", r);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "gr_nmf_create: cannot select device %d", device);
    if (int rc = require_sm100(device)) return rc;

    gr_nmf* h = new (std::nothrow) gr_nmf();
    if (!h) return fail(GR_ERR_OUT_OF_MEMORY, "gr_nmf_create: host allocation failed");
    h->device = device;
    h->n = n;
    h->f = f;
    h->r = r;
    h->rp = r <= 8 ? 8 : (r <= 16 ? 16 : 32);
    // row splits of the W^T X / error kernels: enough CTAs to fill the GPU, >= 256 rows each
    const int slabs = ceil_div(f, kThreads);
    const int64_t want = std::max<int64_t>(1, (148 * 4) / slabs);
    h->splits = (int)std::max<int64_t>(1, std::min<int64_t>(want, ceil_div<int64_t>(n, 256)));
    h->rows_per_split = ceil_div<int64_t>(ceil_div<int64_t>(n, h->splits), 32) * 32;
    h->splits = (int)ceil_div<int64_t>(n, h->rows_per_split);

    auto alloc = [&](auto** p, size_t count) {
        return cudaMalloc(p, count * sizeof(**p)) == cudaSuccess;
    };
    const bool ok = alloc(&h->d_hht, (size_t)r * r) && alloc(&h->d_wtw, (size_t)r * r) &&
                    alloc(&h->d_h_next, (size_t)r * f) &&
                    alloc(&h->d_part_wtx, (size_t)h->splits * h->rp * f) &&
                    alloc(&h->d_part_wtw, (size_t)h->splits * h->rp * r) &&
                    alloc(&h->d_err_part, std::max<size_t>((size_t)slabs * h->splits, kMaxErrParts));
    if (!ok) {
        cudaGetLastError();
        gr_nmf_destroy(h);
        return fail(GR_ERR_OUT_OF_MEMORY, "gr_nmf_create: device allocation failed");
    }
    *out = h;
    return GR_OK;
}

// Zero-padded factors of rank r4 for the tensor-core kernels (r % 4 != 0).  A zero column of W and
// the matching zero row of H are fixed points of the multiplicative updates (numerator 0,
// denominator 0 -> eps, 0 * 0 / eps = 0) and add exact zeros to every sum the real roles see, so
// the first r roles of the padded problem ARE the rank-r problem.
static int ensure_padded(gr_nmf* h) {
    if (h->padded) return GR_OK;
    const int r4 = (h->r + 3) & ~3;
    if (int rc = gr_nmf_create(&h->padded, h->n, h->f, r4, h->device)) return rc;
    if (cudaMalloc(&h->d_wpad, (size_t)h->n * r4 * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&h->d_hpad, (size_t)r4 * h->f * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return fail(GR_ERR_OUT_OF_MEMORY, "gr_nmf: padded factors (rank %d -> %d)", h->r, r4);
    }
    return GR_OK;
}
static int pad_factors(gr_nmf* h, const float* W, const float* H, cudaStream_t st) {
    const int r4 = h->padded->r;
    const int64_t cnt = h->n * r4;
    nmf_pad_w_kernel<<<(unsigned)ceil_div<int64_t>(cnt, 256), 256, 0, st>>>(W, h->n, h->r, r4,
                                                                           h->d_wpad);
    GR_LAUNCH_CHECK("nmf_pad_w_kernel");
    GR_CUDA_TRY(cudaMemcpyAsync(h->d_hpad, H, (size_t)h->r * h->f * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    GR_CUDA_TRY(cudaMemsetAsync(h->d_hpad + (size_t)h->r * h->f, 0,
                                (size_t)(r4 - h->r) * h->f * sizeof(float), st));
    return GR_OK;
}
static int unpad_factors(gr_nmf* h, float* W, float* H, cudaStream_t st) {
    const int64_t cnt = h->n * h->r;
    nmf_unpad_w_kernel<<<(unsigned)ceil_div<int64_t>(cnt, 256), 256, 0, st>>>(
        h->d_wpad, h->n, h->r, h->padded->r, W);
    GR_LAUNCH_CHECK("nmf_unpad_w_kernel");
    GR_CUDA_TRY(cudaMemcpyAsync(H, h->d_hpad, (size_t)h->r * h->f * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    return GR_OK;
}
// would the tensor-core kernels take this X at rank r4?  (asked before anything is allocated)
static bool padded_shape_supported(const gr_nmf* h, const float* X, int64_t ldx) {
    gr_nmf probe;
    probe.n = h->n;
    probe.f = h->f;
    probe.r = (h->r + 3) & ~3;
    probe.device = h->device;
    return nmf_tc_supported(&probe, X, ldx);
}
// true when the call should run on the padded handle (which then exists)
static bool wants_padding(gr_nmf* h, const float* X, int64_t ldx, int* rc) {
    *rc = GR_OK;
    if (h->r % 4 == 0 || getenv("GR_NMF_NO_RANK_PADDING")) return false;
    if (!padded_shape_supported(h, X, ldx)) return false;
    return (*rc = ensure_padded(h)) == GR_OK;
}

extern "C" int gr_nmf_destroy(gr_nmf_t* h) {
    if (!h) return GR_OK;
    DeviceGuard guard(h->device);
    if (h->padded) gr_nmf_destroy(h->padded);
    cudaFree(h->d_wpad);
    cudaFree(h->d_hpad);
    cudaFree(h->d_hht);
    cudaFree(h->d_wtw);
    cudaFree(h->d_h_next);
    cudaFree(h->d_part_wtx);
    cudaFree(h->d_part_wtw);
    cudaFree(h->d_err_part);
    nmf_tc_release(h);
    delete h;
    return GR_OK;
}

extern "C" int gr_nmf_last_path(const gr_nmf_t* h) { return h && h->last_path_tc ? 1 : 0; }

extern "C" int gr_nmf_takes_tensor_cores(const gr_nmf_t* h, const float* X, int64_t ldx) {
    if (!h || !X) return 0;
    if (nmf_tc_supported(h, X, ldx)) return 1;
    if (h->r % 4 == 0 || getenv("GR_NMF_NO_RANK_PADDING")) return 0;
    return padded_shape_supported(h, X, ldx) ? 1 : 0;
}

extern "C" int gr_nmf_error_f32(gr_nmf_t* h, const float* X, int64_t ldx, const float* W,
                                const float* H, double* err_out, void* stream) {
    GR_REQUIRE(h && X && W && H && err_out, "gr_nmf_error_f32: NULL argument");
    GR_REQUIRE(ldx >= h->f, "gr_nmf_error_f32: ldx < f");
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", h->device);
    return nmf_error(h, X, ldx, W, H, err_out, static_cast<cudaStream_t>(stream));
}

extern "C" int gr_nmf_error_tf32(gr_nmf_t* h, const float* X, int64_t ldx, const float* W,
                                 const float* H, double* err_out, void* stream) {
    GR_REQUIRE(h && X && W && H && err_out, "gr_nmf_error_tf32: NULL argument");
    GR_REQUIRE(ldx >= h->f, "gr_nmf_error_tf32: ldx < f");
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", h->device);
    {
        int rc;
        if (wants_padding(h, X, ldx, &rc)) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            if ((rc = pad_factors(h, W, H, st)) != GR_OK) return rc;
            return nmf_error_tc(h->padded, X, ldx, h->d_wpad, h->d_hpad, err_out, st);
        }
        if (rc) return rc;
    }
    GR_REQUIRE(nmf_tc_supported(h, X, ldx),
               "gr_nmf_error_tf32: shape not taken by the tensor-core kernels (r <= 32, f %% 4 == 0, "
               "f <= 1024, 16-byte aligned rows); use gr_nmf_error_f32");
    return nmf_error_tc(h, X, ldx, W, H, err_out, static_cast<cudaStream_t>(stream));
}

// ---- row-sharded form (path B): one iteration cut at the reduction ------
extern "C" int gr_nmf_iteration_local_f32(gr_nmf_t* h, const float* X, int64_t ldx, float* W,
                                          const float* H, int32_t use_tf32, float* wtx_out,
                                          float* wtw_out, void* stream) {
    GR_REQUIRE(h && X && W && H && wtx_out && wtw_out, "gr_nmf_iteration_local_f32: NULL argument");
    GR_REQUIRE(ldx >= h->f, "gr_nmf_iteration_local_f32: ldx < f");
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool tc = use_tf32 && nmf_tc_supported(h, X, ldx);
    h->defer_finish = true;
    // neither path writes H before nmf_finish_iteration, which is deferred here
    const int rc = tc ? nmf_iteration_tc(h, X, ldx, W, const_cast<float*>(H), st)
                      : nmf_iteration_fma(h, X, ldx, W, const_cast<float*>(H), st);
    h->defer_finish = false;
    h->last_path_tc = tc;
    if (rc) return rc;
    dim3 grid((unsigned)ceil_div(h->f, kThreads), (unsigned)h->r);
    nmf_reduce_wtx_kernel<<<grid, kThreads, 0, st>>>(h->pending_wtx, h->pending_splits,
                                                    h->pending_rp, h->f, wtx_out);
    GR_LAUNCH_CHECK("nmf_reduce_wtx_kernel");
    nmf_reduce_wtw_kernel<<<1, 1024, 0, st>>>(h->pending_wtw, h->pending_splits, h->pending_rp,
                                              h->r, wtw_out);
    GR_LAUNCH_CHECK("nmf_reduce_wtw_kernel");
    return GR_OK;
}

extern "C" int gr_nmf_update_h_f32(gr_nmf_t* h, const float* wtx, const float* wtw, float* H,
                                   void* stream) {
    GR_REQUIRE(h && wtx && wtw && H, "gr_nmf_update_h_f32: NULL argument");
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)ceil_div(h->f, kThreads), (unsigned)h->r);
    nmf_update_h_kernel<<<grid, kThreads, 0, st>>>(wtx, 1, h->r, h->r, h->f, wtw, H, h->d_h_next);
    GR_LAUNCH_CHECK("nmf_update_h_kernel");
    GR_CUDA_TRY(cudaMemcpyAsync(H, h->d_h_next, (size_t)h->r * h->f * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    return GR_OK;
}

extern "C" int gr_nmf_mu_f32(gr_nmf_t* h, const float* X, int64_t ldx, float* W, float* H,
                             int32_t max_iter, double tol, int32_t check_every, int32_t use_tf32,
                             int32_t* n_iter_out, double* err_out, void* stream) {
    GR_REQUIRE(h && X && W && H, "gr_nmf_mu_f32: NULL argument");
    GR_REQUIRE(ldx >= h->f, "gr_nmf_mu_f32: ldx = %lld < f = %d", (long long)ldx, h->f);
    GR_REQUIRE(max_iter >= 0 && tol >= 0 && check_every >= 1, "gr_nmf_mu_f32: bad loop control");
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(GR_ERR_CUDA, "cannot select device %d", h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    if (use_tf32) {      // r % 4 != 0: the same loop on zero-padded factors of rank r4
        int rc;
        if (wants_padding(h, X, ldx, &rc)) {
            if ((rc = pad_factors(h, W, H, st)) != GR_OK) return rc;
            rc = gr_nmf_mu_f32(h->padded, X, ldx, h->d_wpad, h->d_hpad, max_iter, tol,
                               check_every, use_tf32, n_iter_out, err_out, stream);
            h->last_path_tc = h->padded->last_path_tc;
            if (rc) return rc;
            return unpad_factors(h, W, H, st);
        }
        if (rc) return rc;
    }
    const bool tc = use_tf32 && nmf_tc_supported(h, X, ldx);
    // the convergence checks follow the path of the iteration: W H on the tensor core beside the
    // tcgen05 iteration kernel (GR_NMF_ERROR_FFMA=1 keeps them on the fp32 FFMA pass)
    static const bool err_ffma = getenv("GR_NMF_ERROR_FFMA") != nullptr;
    auto nmf_error = [&](gr_nmf* hh, const float* Xp, int64_t ld, const float* Wp, const float* Hp,
                         double* e, cudaStream_t s) {
        return tc && !err_ffma ? nmf_error_tc(hh, Xp, ld, Wp, Hp, e, s)
                               : gr::nmf_error(hh, Xp, ld, Wp, Hp, e, s);
    };
    double error_at_init = 0.0, previous = 0.0, error = 0.0;
    if (tol > 0) {
        if (int rc = nmf_error(h, X, ldx, W, H, &error_at_init, st)) return rc;
        previous = error = error_at_init;
    }
    int it = 0;
    for (it = 1; it <= max_iter; ++it) {
        if (int rc = tc ? nmf_iteration_tc(h, X, ldx, W, H, st)
                        : nmf_iteration_fma(h, X, ldx, W, H, st))
            return rc;
        if (tol > 0 && it % check_every == 0) {
            if (int rc = nmf_error(h, X, ldx, W, H, &error, st)) return rc;
            if ((previous - error) / error_at_init < tol) break;   // _nmf.py:877
            previous = error;
        }
    }
    if (it > max_iter) it = max_iter;
    if (tol == 0 && err_out)   // no convergence test ran: report the final error
        if (int rc = nmf_error(h, X, ldx, W, H, &error, st)) return rc;
    if (n_iter_out) *n_iter_out = it;
    if (err_out) *err_out = error;
    h->last_path_tc = tc;
    return GR_OK;
}
