// State block of the DagmaLinear inner iteration on the multi-CTA path (shared by large.cu and lin_iter.cu; mirrored by
// midagma_b200/_large.py).
#pragma once
#include <cstdint>

namespace dagma {

struct LinState {           // mirrored by midagma_b200/_large.py (all 8-byte fields first)
    double mu, s, lr, lambda1, beta1, beta2;
    double p1_hi, p1_lo, p2_hi, p2_lo;      // beta^it as double-double
    double logabsdet, h, min_entry;
    double score_acc, l1_acc, loss_acc;     // reduction outputs
    double gscale;                          // l2: 1 (T = cov W), logistic: 1/n (T = X^T sigmoid(XW))
    int32_t it, halted, info, pad;
};

// (hi, lo) *= b in double-double arithmetic: beta^it stays exact to ~1e-32 over 1e5 iterations
__device__ __forceinline__ void dd_mul(double& hi, double& lo, double b) {
    const double ph = hi * b;
    const double pl = fma(hi, b, -ph) + lo * b;
    const double s = ph + pl;
    lo = pl - (s - ph);
    hi = s;
}

}  // namespace dagma
