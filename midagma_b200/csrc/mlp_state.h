// State block of the DagmaMLP inner iteration (shared by mlp.cu and mlp_iter.cu; mirrored by midagma_b200/nonlinear.py).
#pragma once
#include <cstdint>

namespace dagma {

struct MlpState {            // mirrored by midagma_b200/nonlinear.py
    double mu, s, lr, lambda1, lambda2, beta1, beta2;
    double logabsdet, h, min_entry;       // written by the logdet kernel
    double S, l1, obj, score;             // S = sum res^2 (this rank, then global), l1 = sum |W1|
    double lr_gamma;                      // ExponentialLR factor applied every 1000 steps (1 = off)
    int32_t step, halted, info, pad;
};

}  // namespace dagma
