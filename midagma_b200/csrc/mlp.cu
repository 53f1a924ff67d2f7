// DagmaMLP / DagmaNonlinear inner iteration for dims = [d, m1, 1] (reference:
// src/dagma/nonlinear.py:45-97, 139-159, 208-225; src/dagma/locally_connected.py:55-75).
//
// Parameters live in one flat FP64 buffer  theta = [W1 (P x d) | b1 (P) | W2 (P) | b2 (d)],
// P = d*m1, row p = j*m1 + k of W1 is hidden unit k of node j (fc1.weight layout).
// Activations are kept TRANSPOSED ([P][n], sample index fastest) so that both GEMMs are
// plain row-major products on the DMMA kernel and every element-wise pass is coalesced:
//     Zt = W1 Xt            (P x n,  K = d)
//     gW1 = dZt X           (P x d,  K = n, split-K)
// The log couples all samples only through S = sum res^2, so all sums are accumulated
// un-scaled (d/S is applied in the Adam kernel): with rows of X sharded over GPUs one
// all-reduce of [grads | S] per iteration is enough.
#include "common.cuh"
#include "small_gj.cuh"
#include "mlp_state.h"
#include "../../include/dagma_b200.h"

namespace dagma {

// A[i][j] = sum_k W1[(j*m1+k)][i]^2  (nonlinear.py:82-84)  + partial sums of |W1|
__global__ void mlp_adj_kernel(const double* __restrict__ W1, int d, int m1, double* __restrict__ A,
                               double* __restrict__ l1_partial) {
    __shared__ double red[96];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;      // e = j*d + i  (coalesced over i)
    double a = 0.0, l1 = 0.0, z = 0.0;
    if (e < d * d) {
        const int j = e / d, i = e - j * d;
        for (int k = 0; k < m1; ++k) {
            const double w = W1[(size_t)(j * m1 + k) * d + i];
            a = fma(w, w, a);
            l1 += fabs(w);
        }
        A[(size_t)i * d + j] = a;
    }
    double dummy = 0.0;
    block_sum3<256>(l1, z, dummy, red, threadIdx.x);
    if (threadIdx.x == 0) l1_partial[blockIdx.x] = l1;
}

// thread 0 of every block, after the block's partial results are written: true in the LAST block to get here, which has
// then seen the partial results of all the others (fence / atomic / fence) -- it finishes the reduction in a fixed
// order, so the result does not depend on which block that is and no separate reduction launch is needed
__device__ __forceinline__ bool mlp_last_block(MlpState* st, int blocks) {
    __threadfence();
    const bool last = (atomicAdd(&st->pad, 1) == blocks - 1);
    if (last) {
        st->pad = 0;
        __threadfence();
    }
    return last;
}
// S = sum res^2 and l1 = sum |W1| -> state (fixed order)
__device__ __forceinline__ void mlp_finish_forward(MlpState* st, const double* S_partial, int nS, const double* l1_partial,
                                                   int nl1) {
    double S = 0.0, l1 = 0.0;
    for (int i = 0; i < nS; ++i) S += __ldcg(S_partial + i);
    for (int i = 0; i < nl1; ++i) l1 += __ldcg(l1_partial + i);
    st->S = S;
    st->l1 = l1;
}
// forward tail: H = sigmoid(Zt + b1) (in place), out = sum_k H W2 + b2, res = out - Xt, partial S
__global__ void __launch_bounds__(256) mlp_forward_kernel(double* __restrict__ Zt, const double* __restrict__ theta,
                                                          const double* __restrict__ Xt, int n, int d, int m1,
                                                          double* __restrict__ res, double* __restrict__ out_opt,
                                                          double* __restrict__ S_partial, MlpState* st,
                                                          const double* l1_partial, int nl1) {
    __shared__ double red[96];
    const int P = d * m1;
    const double* b1 = theta + (size_t)P * d;
    const double* W2 = b1 + P;
    const double* b2 = W2 + P;
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0, z1 = 0.0, z2 = 0.0;
    if (s < n) {
        double o = b2[j];
        for (int k = 0; k < m1; ++k) {
            const int p = j * m1 + k;
            const double zz = Zt[(size_t)p * n + s] + b1[p];
            const double hh = 1.0 / (1.0 + exp(-zz));
            Zt[(size_t)p * n + s] = hh;
            o = fma(hh, W2[p], o);
        }
        const double r = o - Xt[(size_t)j * n + s];
        res[(size_t)j * n + s] = r;
        if (out_opt) out_opt[(size_t)s * d + j] = o;
        sq = r * r;
    }
    block_sum3<256>(sq, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) {
        S_partial[blockIdx.y * gridDim.x + blockIdx.x] = sq;
        if (mlp_last_block(st, gridDim.x * gridDim.y)) mlp_finish_forward(st, S_partial, gridDim.x * gridDim.y, l1_partial, nl1);
    }
}

// after the (optional) all-reduce of S
__global__ void mlp_objective_kernel(MlpState* st, int n_total, int d) {
    if (threadIdx.x != 0) return;
    const double score = 0.5 * (double)d * log(st->S / (double)n_total);
    st->score = score;
    st->obj = st->mu * (score + st->lambda1 * st->l1) + st->h;
    if (st->halted == 0 && st->h < 0.0) st->halted = 1;
}

// backward tail: dZt (un-scaled, overwrites H) and the partial sums over this block's samples of
//   gW2[p] = sum res H, gb1[p] = sum dZ, gb2[j] = sum res      (SURVEY.md "validated restatements")
__global__ void __launch_bounds__(256) mlp_backward_kernel(double* __restrict__ Ht, const double* __restrict__ theta,
                                                           const double* __restrict__ res, int n, int d, int m1,
                                                           double* __restrict__ part) {   // [chunks][2P + d]
    __shared__ double red[96];
    const int P = d * m1;
    const double* W2 = theta + (size_t)P * d + P;
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = s < n;
    const double r = ok ? res[(size_t)j * n + s] : 0.0;
    double* row = part + (size_t)blockIdx.x * (2 * P + d);
    for (int k = 0; k < m1; ++k) {
        const int p = j * m1 + k;
        double gw2 = 0.0, gb1 = 0.0, z = 0.0;
        if (ok) {
            const double hh = Ht[(size_t)p * n + s];
            const double dz = r * W2[p] * hh * (1.0 - hh);
            Ht[(size_t)p * n + s] = dz;
            gw2 = r * hh;
            gb1 = dz;
        }
        block_sum3<256>(gw2, gb1, z, red, threadIdx.x);
        if (threadIdx.x == 0) {
            row[P + p] = gw2;          // layout of `part` row: [gb1 (P) | gW2 (P) | gb2 (d)]
            row[p] = gb1;
        }
    }
    double rr = r, z1 = 0.0, z2 = 0.0;
    block_sum3<256>(rr, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) row[2 * P + j] = rr;
}
// grads[P*d ...] = sum over chunks (fixed order)
__global__ void mlp_reduce_parts_kernel(const double* __restrict__ part, int chunks, int width, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= width) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * width + e];
    out[e] = s;
}

// end of an iteration: step counter and ExponentialLR (nonlinear.py:224-225).  Called by thread 0 of every block of the
// Adam kernel when the block is done; the LAST block to get here has seen all others finish (nobody reads the state
// any more) and advances it -- no separate 1-thread launch.
__device__ __forceinline__ void mlp_advance_by_last_block(MlpState* st, int blocks) {
    __threadfence();
    if (atomicAdd(&st->pad, 1) == blocks - 1) {
        st->pad = 0;
        if (st->halted == 0) {
            st->step += 1;
            if (st->lr_gamma != 1.0 && (st->step % 1000) == 0) st->lr *= st->lr_gamma;
        }
    }
}

// torch.optim.Adam step (single-tensor form, weight decay added to the gradient) over theta.
// grads hold UN-SCALED score sums; the factor d/S and mu are applied here; fc1.weight also gets
// mu*lambda1*sign(w) and dh/dw = 2 w Minv[j][i]   (nonlinear.py:208, 220-223).
__global__ void __launch_bounds__(256) mlp_adam_kernel(const MlpState* st, double* __restrict__ theta,
                                                       const double* __restrict__ grads, double* __restrict__ m,
                                                       double* __restrict__ v, const double* __restrict__ Minv,
                                                       int d, int m1, size_t total) {
    if (st->halted != 0) {                    // uniform over the grid
        if (threadIdx.x == 0) mlp_advance_by_last_block(const_cast<MlpState*>(st), gridDim.x);
        return;
    }
    const double mu = st->mu, lr = st->lr, b1 = st->beta1, b2 = st->beta2;
    const double wd = mu * st->lambda2, l1c = mu * st->lambda1;
    const double gs = mu * (double)d / st->S;
    const int step = st->step + 1;
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    const double step_size = lr / bc1, bc2s = sqrt(bc2);
    const size_t nW1 = (size_t)d * m1 * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const double p = theta[e];
        double g = gs * grads[e];
        if (e < nW1) {
            const int row = (int)(e / d), i = (int)(e - (size_t)row * d), j = row / m1;
            const double sg = (p > 0.0) ? 1.0 : ((p < 0.0) ? -1.0 : 0.0);
            g = fma(l1c, sg, g);
            g = fma(2.0 * p, Minv[(size_t)j * d + i], g);
        }
        g = fma(wd, p, g);
        const double mn = m[e] + (g - m[e]) * (1.0 - b1);            // exp_avg.lerp_(grad, 1 - beta1)
        const double vn = fma(v[e], b2, (1.0 - b2) * g * g);
        m[e] = mn;
        v[e] = vn;
        const double denom = sqrt(vn) / bc2s + 1e-8;
        theta[e] = p - step_size * (mn / denom);
    }
    __syncthreads();                           // every thread of the block has read the state
    if (threadIdx.x == 0) mlp_advance_by_last_block(const_cast<MlpState*>(st), gridDim.x);
}

}  // namespace dagma

using namespace dagma;

extern "C" int dagma_mlp_adj_f64(dagma_stream_t stream, int d, int m1, const double* theta_dev, double* a_dev,
                                 double* l1_partial_dev) {
    DAGMA_REQUIRE(theta_dev && a_dev && l1_partial_dev, "null pointer");
    mlp_adj_kernel<<<(d * d + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta_dev, d, m1, a_dev, l1_partial_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_forward_f64(dagma_stream_t stream, int n, int d, int m1, double* zt_dev,
                                     const double* theta_dev, const double* xt_dev, double* res_dev,
                                     double* out_opt_dev, double* s_partial_dev, const double* l1_partial_dev,
                                     void* state_dev) {
    DAGMA_REQUIRE(zt_dev && theta_dev && xt_dev && res_dev && s_partial_dev && state_dev, "null pointer");
    dim3 grid((n + 255) / 256, d);
    mlp_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(zt_dev, theta_dev, xt_dev, n, d, m1, res_dev, out_opt_dev,
                                                              s_partial_dev, (MlpState*)state_dev, l1_partial_dev,
                                                              l1_partial_dev ? (d * d + 255) / 256 : 0);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_objective_f64(dagma_stream_t stream, void* state_dev, int n_total, int d) {
    mlp_objective_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((MlpState*)state_dev, n_total, d);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_backward_f64(dagma_stream_t stream, int n, int d, int m1, double* ht_dev,
                                      const double* theta_dev, const double* res_dev, double* part_dev,
                                      double* grads_tail_dev) {
    DAGMA_REQUIRE(ht_dev && theta_dev && res_dev && part_dev && grads_tail_dev, "null pointer");
    const int chunks = (n + 255) / 256, width = 2 * d * m1 + d;
    dim3 grid(chunks, d);
    mlp_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ht_dev, theta_dev, res_dev, n, d, m1, part_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    mlp_reduce_parts_kernel<<<(width + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part_dev, chunks, width, grads_tail_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_adam_f64(dagma_stream_t stream, int d, int m1, void* state_dev, double* theta_dev,
                                  const double* grads_dev, double* m_dev, double* v_dev, const double* minv_dev) {
    DAGMA_REQUIRE(state_dev && theta_dev && grads_dev && m_dev && v_dev && minv_dev, "null pointer");
    const size_t total = (size_t)d * m1 * d + 2 * (size_t)d * m1 + d;
    const int blocks = (int)((total + 255) / 256);
    mlp_adam_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, (cudaStream_t)stream>>>((const MlpState*)state_dev, theta_dev,
                                                                                     grads_dev, m_dev, v_dev, minv_dev, d, m1, total);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- general LocallyConnected stacks  dims = [d, m_1, ..., m_L, 1]  (nonlinear.py:39-43, 60-65) ----------
// Same transposed activations ([d * width][n]) and un-scaled sums as the fused [d, m1, 1] kernels above; one
// forward and one backward launch per layer.  theta = [W1 | b1 | W_0 | b_0 | ... | W_{L-1} | b_{L-1}] with
// W_l = fc2.l.weight [d][mi][mo] and b_l = fc2.l.bias [d][mo].
namespace dagma {
// in: pre-activations of the layer's input ([d*mi][n]; + bias_in[d*mi] for the first layer), replaced by
// H = sigmoid(.) for the backward pass; out[(j*mo+o)][s] = b[j][o] + sum_k H[j*mi+k][s] W[j][k][o]
__global__ void __launch_bounds__(256) lc_forward_kernel(double* __restrict__ in, const double* __restrict__ bias_in,
                                                         const double* __restrict__ W, const double* __restrict__ b,
                                                         double* __restrict__ out, int n, int d, int mi, int mo) {
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    for (int k = 0; k < mi; ++k) {
        const int p = j * mi + k;
        const double zz = in[(size_t)p * n + s] + (bias_in ? bias_in[p] : 0.0);
        in[(size_t)p * n + s] = 1.0 / (1.0 + exp(-zz));
    }
    for (int o = 0; o < mo; ++o) {
        double acc = b ? b[j * mo + o] : 0.0;
        for (int k = 0; k < mi; ++k)
            acc = fma(in[(size_t)(j * mi + k) * n + s], W[((size_t)j * mi + k) * mo + o], acc);
        out[(size_t)(j * mo + o) * n + s] = acc;
    }
}
// res = out (+ bias) - Xt, partial sums of res^2 (one per block); optional row-major copy of the prediction
__global__ void __launch_bounds__(256) mlp_residual_kernel(const double* __restrict__ out, const double* __restrict__ bias,
                                                           const double* __restrict__ Xt, int n, int d,
                                                           double* __restrict__ res, double* __restrict__ out_opt,
                                                           double* __restrict__ S_partial, MlpState* st,
                                                          const double* l1_partial, int nl1) {
    __shared__ double red[96];
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0, z1 = 0.0, z2 = 0.0;
    if (s < n) {
        const double o = out[(size_t)j * n + s] + (bias ? bias[j] : 0.0);
        const double r = o - Xt[(size_t)j * n + s];
        res[(size_t)j * n + s] = r;
        if (out_opt) out_opt[(size_t)s * d + j] = o;
        sq = r * r;
    }
    block_sum3<256>(sq, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) {
        S_partial[blockIdx.y * gridDim.x + blockIdx.x] = sq;
        if (mlp_last_block(st, gridDim.x * gridDim.y)) mlp_finish_forward(st, S_partial, gridDim.x * gridDim.y, l1_partial, nl1);
    }
}
// H ([d*mi][n], from lc_forward_kernel) is replaced by dZ = (sum_o W[j][k][o] dZn[j*mo+o]) H (1 - H);
// part row of this sample chunk: [gW (d*mi*mo) | gb (d*mo)] = sums over the chunk of H dZn and dZn
__global__ void __launch_bounds__(256) lc_backward_kernel(double* __restrict__ H, const double* __restrict__ dZn,
                                                          const double* __restrict__ W, int n, int d, int mi, int mo,
                                                          double* __restrict__ part) {
    __shared__ double red[96];
    const int j = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = s < n;
    double* row = part + (size_t)blockIdx.x * ((size_t)d * mi * mo + (size_t)d * mo);
    for (int k = 0; k < mi; ++k) {
        const int p = j * mi + k;
        const double hh = ok ? H[(size_t)p * n + s] : 0.0;
        double dh = 0.0;
        for (int o = 0; o < mo; o += 3) {
            double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (o + q < mo && ok) {
                    const double dz = dZn[(size_t)(j * mo + o + q) * n + s];
                    g[q] = hh * dz;
                    dh = fma(W[((size_t)j * mi + k) * mo + o + q], dz, dh);
                }
            block_sum3<256>(g[0], g[1], g[2], red, threadIdx.x);
            if (threadIdx.x == 0)
                for (int q = 0; q < 3 && o + q < mo; ++q) row[((size_t)j * mi + k) * mo + o + q] = g[q];
        }
        if (ok) H[(size_t)p * n + s] = dh * hh * (1.0 - hh);
    }
    for (int o = 0; o < mo; o += 3) {
        double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (o + q < mo && ok) g[q] = dZn[(size_t)(j * mo + o + q) * n + s];
        block_sum3<256>(g[0], g[1], g[2], red, threadIdx.x);
        if (threadIdx.x == 0)
            for (int q = 0; q < 3 && o + q < mo; ++q) row[(size_t)d * mi * mo + (size_t)j * mo + o + q] = g[q];
    }
}
// out[r] = sum_s a[r][s]   (one block per row, fixed order)
__global__ void __launch_bounds__(256) row_sums_kernel(const double* __restrict__ a, int n, double* __restrict__ out) {
    __shared__ double red[96];
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    const double* row = a + (size_t)blockIdx.x * n;
    for (int s = threadIdx.x; s < n; s += 256) acc += row[s];
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}
}  // namespace dagma

extern "C" int dagma_lc_forward_f64(dagma_stream_t stream, int n, int d, int mi, int mo, double* in_dev,
                                    const double* bias_in_dev, const double* w_dev, const double* b_dev,
                                    double* out_dev) {
    DAGMA_REQUIRE(in_dev && w_dev && out_dev && n >= 1 && d >= 1 && mi >= 1 && mo >= 1, "bad arguments");
    dim3 grid((n + 255) / 256, d);
    dagma::lc_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in_dev, bias_in_dev, w_dev, b_dev, out_dev, n, d, mi, mo);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_residual_f64(dagma_stream_t stream, int n, int d, const double* out_dev,
                                      const double* bias_dev, const double* xt_dev, double* res_dev,
                                      double* out_opt_dev, double* s_partial_dev, const double* l1_partial_dev,
                                      void* state_dev) {
    DAGMA_REQUIRE(out_dev && xt_dev && res_dev && s_partial_dev && state_dev, "null pointer");
    dim3 grid((n + 255) / 256, d);
    dagma::mlp_residual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out_dev, bias_dev, xt_dev, n, d, res_dev, out_opt_dev,
                                                                      s_partial_dev, (MlpState*)state_dev, l1_partial_dev,
                                                                      l1_partial_dev ? (d * d + 255) / 256 : 0);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_lc_backward_f64(dagma_stream_t stream, int n, int d, int mi, int mo, double* h_dev,
                                     const double* dzn_dev, const double* w_dev, double* part_dev,
                                     double* grads_dev) {
    DAGMA_REQUIRE(h_dev && dzn_dev && w_dev && part_dev && grads_dev, "null pointer");
    const int chunks = (n + 255) / 256, width = d * mi * mo + d * mo;
    dim3 grid(chunks, d);
    dagma::lc_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h_dev, dzn_dev, w_dev, n, d, mi, mo, part_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    mlp_reduce_parts_kernel<<<(width + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part_dev, chunks, width, grads_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_row_sums_f64(dagma_stream_t stream, int rows, int n, const double* a_dev, double* out_dev) {
    DAGMA_REQUIRE(a_dev && out_dev && rows >= 1 && n >= 1, "bad arguments");
    dagma::row_sums_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(a_dev, n, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_adam_ex_f64(dagma_stream_t stream, int d, int m1, size_t total, void* state_dev,
                                     double* theta_dev, const double* grads_dev, double* m_dev, double* v_dev,
                                     const double* minv_dev) {
    DAGMA_REQUIRE(state_dev && theta_dev && grads_dev && m_dev && v_dev && minv_dev, "null pointer");
    DAGMA_REQUIRE(total >= (size_t)d * m1 * d, "total is smaller than fc1.weight");
    const int blocks = (int)((total + 255) / 256);
    mlp_adam_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, (cudaStream_t)stream>>>((const MlpState*)state_dev, theta_dev,
                                                                                     grads_dev, m_dev, v_dev, minv_dev, d, m1, total);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- stand-alone helpers of the public surface -------------------------------------------
namespace dagma {
// LocallyConnected.forward (locally_connected.py:70-74): out[n,j,:] = in[n,j,:] @ W[j] + b[j]
__global__ void locally_connected_kernel(const double* __restrict__ in, const double* __restrict__ W,
                                         const double* __restrict__ b, double* __restrict__ out, int n, int d,
                                         int m1, int m2) {
    const size_t total = (size_t)n * d * m2;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(e % m2);
        const int j = (int)((e / m2) % d);
        const size_t s = e / ((size_t)m2 * d);
        double acc = b ? b[(size_t)j * m2 + o] : 0.0;
        const double* x = in + (s * d + j) * m1;
        const double* w = W + (size_t)j * m1 * m2 + o;
        for (int k = 0; k < m1; ++k) acc = fma(x[k], w[(size_t)k * m2], acc);
        out[e] = acc;
    }
}
__global__ void __launch_bounds__(256) sumsq_diff_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                         size_t total, double* partial) {
    __shared__ double red[96];
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const double dlt = a[e] - b[e];
        acc = fma(dlt, dlt, acc);
    }
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void sum_small_kernel(const double* partial, int n, double* out) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = s;
}
}  // namespace dagma

extern "C" int dagma_locally_connected_f64(dagma_stream_t stream, int n, int d, int m1, int m2, const double* in_dev,
                                           const double* w_dev, const double* b_dev, double* out_dev) {
    DAGMA_REQUIRE(in_dev && w_dev && out_dev, "null pointer");
    const size_t total = (size_t)n * d * m2;
    const int blocks = (int)((total + 255) / 256);
    dagma::locally_connected_kernel<<<blocks < 2368 ? blocks : 2368, 256, 0, (cudaStream_t)stream>>>(in_dev, w_dev, b_dev,
                                                                                                   out_dev, n, d, m1, m2);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_sumsq_diff_f64(dagma_stream_t stream, size_t total, const double* a_dev, const double* b_dev,
                                    double* partial_dev, int n_partial, double* out_dev) {
    DAGMA_REQUIRE(a_dev && b_dev && partial_dev && out_dev && n_partial >= 1, "bad arguments");
    const int blocks = n_partial < 592 ? n_partial : 592;
    dagma::sumsq_diff_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a_dev, b_dev, total, partial_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    dagma::sum_small_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(partial_dev, blocks, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- fork's TCC log-det constraint: block assembly and fold-back (src/notreks/notreks.py) ----
namespace dagma {
// A = [[W o W, w S], [I, (W o W)^T]]  (n = 2d);  with_S = 0 builds the baseline B (zero block)
__global__ void tcc_assemble_kernel(const double* __restrict__ W, const double* __restrict__ S, double w, int d,
                                    int with_S, double* __restrict__ A) {
    const int n = 2 * d;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    const int r = e / n, c = e - r * n;
    double val;
    if (r < d && c < d) { const double x = W[(size_t)r * d + c]; val = x * x; }
    else if (r < d) val = with_S ? w * S[(size_t)r * d + (c - d)] : 0.0;
    else if (c < d) val = (r - d == c) ? 1.0 : 0.0;
    else { const double x = W[(size_t)(c - d) * d + (r - d)]; val = x * x; }
    A[e] = val;
}
// out (+)= sign * 2 W o (G11 + G22^T)    (notreks.py:285-287, 384)
__global__ void tcc_fold_kernel(const double* __restrict__ W, const double* __restrict__ G, int d, double sign,
                                int accumulate, double* __restrict__ out) {
    const int n = 2 * d;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d * d) return;
    const int r = e / d, c = e - r * d;
    const double g = G[(size_t)r * n + c] + G[(size_t)(d + c) * n + (d + r)];
    const double val = sign * 2.0 * W[e] * g;
    out[e] = accumulate ? out[e] + val : val;
}
}  // namespace dagma

extern "C" int dagma_tcc_assemble_f64(dagma_stream_t stream, int d, const double* w_dev, const double* s_dev,
                                      double w, int with_s, double* a_dev) {
    DAGMA_REQUIRE(w_dev && a_dev && (s_dev || !with_s), "null pointer");
    dagma::tcc_assemble_kernel<<<(4 * d * d + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_dev, s_dev, w, d, with_s, a_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int dagma_tcc_fold_f64(dagma_stream_t stream, int d, const double* w_dev, const double* g_dev, double sign,
                                  int accumulate, double* out_dev) {
    DAGMA_REQUIRE(w_dev && g_dev && out_dev, "null pointer");
    dagma::tcc_fold_kernel<<<(d * d + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_dev, g_dev, d, sign, accumulate, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
