// Batched evaluation of estimated graphs against the truth (utils.count_accuracy / utils.is_dag of the
// reference, src/dagma/utils.py:13-18, 245-310, without igraph): the integer counts behind fdr / tpr / fpr / shd / nnz and
// the acyclicity test for a whole batch of problems in one launch -- a sweep of thousands of fits is scored where its
// results already are.  One CTA per problem: a coalesced pass over the d x d entries for the counts and the in-degrees
// (the transposed entries come through L2), then Kahn's algorithm by rounds -- every round removes ALL current sources,
// so the work is O(d^2) in total and the number of block barriers is the length of the longest path.
#include "common.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

constexpr int GM_NT = 256;

// counts[p][0..7]: nnz (prediction positive, undirected ones included), condition positive, true positive, false
// positive, reverse, extra (lower triangle of the skeleton), missing (lower), is_dag
__global__ void __launch_bounds__(GM_NT) graph_metrics_kernel(int batch, int d, const int8_t* __restrict__ est,
                                                              const uint8_t* __restrict__ tru, int true_shared,
                                                              int* __restrict__ counts) {
    extern __shared__ int gsm[];
    int* indeg = gsm;                 // [d] in-degree over the nodes still alive (directed edges of the estimate)
    int* state = gsm + d;             // [d] 0 alive, 1 leaving in this round, 2 gone
    __shared__ int s_cnt[8];
    __shared__ int s_left, s_moved;
    const int tid = threadIdx.x;
    for (int p = blockIdx.x; p < batch; p += gridDim.x) {
        const int8_t* E = est + (size_t)p * d * d;
        const uint8_t* T = tru + (true_shared ? 0 : (size_t)p * d * d);
        for (int i = tid; i < d; i += GM_NT) { indeg[i] = 0; state[i] = 0; }
        if (tid < 8) s_cnt[tid] = 0;
        __syncthreads();
        int c_pred = 0, c_cond = 0, c_tp = 0, c_fp = 0, c_rev = 0, c_extra = 0, c_miss = 0;
        for (int r = tid / 32 + 0; r < d; r += GM_NT / 32)            // a warp per row: coalesced over c
            for (int c = tid & 31; c < d; c += 32) {
                const int e = E[(size_t)r * d + c], et = E[(size_t)c * d + r];
                const bool t = T[(size_t)r * d + c] != 0, tt = T[(size_t)c * d + r] != 0;
                const bool dir = (e == 1), und = (e == -1), skel = t || tt;
                c_pred += dir || und;
                c_cond += t;
                c_tp += (dir && t) || (und && skel);
                c_fp += (dir || und) && !skel;
                c_rev += dir && !t && tt;
                if (r >= c) {                                        // np.tril(B + B.T) != 0, literally
                    const bool pl = (e + et) != 0, cl = t || tt;
                    c_extra += pl && !cl;
                    c_miss += cl && !pl;
                }
                if (e != 0) atomicAdd(&indeg[c], 1);
            }
        // warp sums, then one shared atomic per warp and counter (integers: any order gives the same result)
        int v[7] = {c_pred, c_cond, c_tp, c_fp, c_rev, c_extra, c_miss};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            int x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if ((tid & 31) == 0 && x) atomicAdd(&s_cnt[k], x);
        }
        if (tid == 0) s_left = d;
        __syncthreads();
        // Kahn by rounds
        for (;;) {
            if (tid == 0) s_moved = 0;
            __syncthreads();
            int moved = 0;
            for (int i = tid; i < d; i += GM_NT)
                if (state[i] == 0 && indeg[i] == 0) { state[i] = 1; ++moved; }
            if (moved) atomicAdd(&s_moved, moved);
            __syncthreads();
            const int m = s_moved;
            if (m == 0) break;
            // the out-edges of the leaving nodes: a warp per leaving row
            for (int r = tid / 32; r < d; r += GM_NT / 32) {
                if (state[r] != 1) continue;
                for (int c = tid & 31; c < d; c += 32)
                    if (E[(size_t)r * d + c] != 0) atomicSub(&indeg[c], 1);
            }
            __syncthreads();
            for (int i = tid; i < d; i += GM_NT)
                if (state[i] == 1) state[i] = 2;
            if (tid == 0) s_left -= m;
            __syncthreads();
        }
        if (tid < 7) counts[(size_t)p * 8 + tid] = s_cnt[tid];
        if (tid == 7) counts[(size_t)p * 8 + 7] = (s_left == 0) ? 1 : 0;
        __syncthreads();
    }
}

}  // namespace dagma

using namespace dagma;

extern "C" int dagma_graph_metrics(dagma_stream_t stream, int batch, int d, const int8_t* est_dev, const uint8_t* true_dev,
                                   int true_shared, int* counts_dev) {
    DAGMA_REQUIRE(batch >= 1 && d >= 1 && est_dev && true_dev && counts_dev, "bad arguments");
    const size_t smem = 2 * (size_t)d * sizeof(int);
    DAGMA_REQUIRE(smem <= 200 * 1024, "d too large for the on-chip in-degree arrays");
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        DAGMA_CUDA_OK(cudaFuncSetAttribute(graph_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const int grid = batch < 4 * 148 ? batch : 4 * 148;
    graph_metrics_kernel<<<grid, GM_NT, smem, (cudaStream_t)stream>>>(batch, d, est_dev, true_dev, true_shared, counts_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
