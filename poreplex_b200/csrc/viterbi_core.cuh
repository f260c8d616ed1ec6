// viterbi_core.cuh -- device pieces of the log-space Viterbi shared by the segmentation
// kernel and the unsplit-read (chimera) kernels.  See kernels_viterbi.cu for the
// reference citations and the arithmetic contract.
#pragma once
#include "pb_internal.h"
#include "pb_math.cuh"

namespace pb {

struct HmmMask {
    // has_edge[dst] bit src ; has_start bit s
    uint32_t edge[PB2_MAX_STATES];
    double logp[PB2_MAX_STATES][PB2_MAX_STATES];   // [dst][src]
};

__device__ __forceinline__ void hmm_emissions(const HmmDev &M, double x,
                                              double (&e)[PB2_MAX_STATES])
{
#pragma unroll
    for (int s = 0; s < PB2_MAX_STATES; s++) {
        e[s] = 0.0;
        if (s < M.n_states) {
            const int nc = M.n_comp[s];
            if (nc == 1) {
                const double d = pb::dsub(x, M.mu[s][0]);
                e[s] = pb::dsub(M.log_norm[s][0], pb::dmul(pb::dmul(d, d), M.inv_two_var[s][0]));
            } else {
                double acc = pb::neg_inf();
#pragma unroll
                for (int j = 0; j < PB2_MAX_COMP; j++) {
                    if (j < nc) {
                        const double d = pb::dsub(x, M.mu[s][j]);
                        const double lp = pb::dsub(M.log_norm[s][j],
                                                   pb::dmul(pb::dmul(d, d), M.inv_two_var[s][j]));
                        acc = pb::pair_lse(acc, pb::dadd(lp, M.log_weight[s][j]));
                    }
                }
                e[s] = acc;
            }
        }
    }
}

// Emissions with the number of mixture components per state known at compile time (NCOMP: 4
// bits per state).  pair_lse(-inf, a) returns a, so a two-component state is pair_lse(a0, a1):
// the same values as the generic loop, without its per-component tests.
template <uint32_t NCOMP, int NS>
__device__ __forceinline__ void hmm_emissions_topo(const HmmDev &M, double x,
                                                   double (&e)[PB2_MAX_STATES])
{
#pragma unroll
    for (int s = 0; s < PB2_MAX_STATES; s++) {
        e[s] = 0.0;
        if (s < NS) {
            const int nc = (NCOMP >> (4 * s)) & 15;
            double lp[PB2_MAX_COMP];
#pragma unroll
            for (int j = 0; j < PB2_MAX_COMP; j++) {
                if (j < nc) {
                    const double d = pb::dsub(x, M.mu[s][j]);
                    lp[j] = pb::dsub(M.log_norm[s][j], pb::dmul(pb::dmul(d, d), M.inv_two_var[s][j]));
                }
            }
            if (nc == 1) {
                e[s] = lp[0];
            } else {
                double acc = pb::dadd(lp[0], M.log_weight[s][0]);
#pragma unroll
                for (int j = 1; j < PB2_MAX_COMP; j++)
                    if (j < nc) acc = pb::pair_lse(acc, pb::dadd(lp[j], M.log_weight[s][j]));
                e[s] = acc;
            }
        }
    }
}

inline uint32_t pack_ncomp(const HmmDev &M)
{
    uint32_t m = 0;
    for (int s = 0; s < PB2_MAX_STATES; s++)
        if (s < M.n_states) m |= (uint32_t)(M.n_comp[s] & 15) << (4 * s);
    return m;
}

// One Viterbi time step.  Returns the packed back-pointer word (3 bits per state,
// 7 = no predecessor).
__device__ __forceinline__ uint32_t viterbi_step(const HmmDev &M, const HmmMask &K,
                                                 double (&v)[PB2_MAX_STATES],
                                                 const double (&e)[PB2_MAX_STATES])
{
    double nv[PB2_MAX_STATES];
    uint32_t bp = 0;
#pragma unroll
    for (int l = 0; l < PB2_MAX_STATES; l++) {
        double best = pb::neg_inf();
        uint32_t arg = 7;
        if (l < M.n_states) {
#pragma unroll
            for (int src = 0; src < PB2_MAX_STATES; src++) {
                if (K.edge[l] & (1u << src)) {
                    const double cand = pb::dadd(pb::dadd(v[src], K.logp[l][src]), e[l]);
                    if (cand > best) { best = cand; arg = src; }
                }
            }
        }
        nv[l] = best;
        bp |= arg << (3 * l);
    }
#pragma unroll
    for (int l = 0; l < PB2_MAX_STATES; l++) v[l] = nv[l];
    return bp;
}

// Same step with the topology known at compile time: EDGES packs the in-edge bit masks of the
// eight destination states (8 bits each, bit src of byte dst).  Only the existing edges are
// compiled -- the generic version tests all 64 (dst, src) pairs every step -- in the same
// ascending-src order with the same strict '>', so results are bit-identical.
template <uint64_t EDGES, int NS>
__device__ __forceinline__ uint32_t viterbi_step_topo(const HmmMask &K, double (&v)[PB2_MAX_STATES],
                                                      const double (&e)[PB2_MAX_STATES])
{
    double nv[PB2_MAX_STATES];
    uint32_t bp = 0;
#pragma unroll
    for (int l = 0; l < PB2_MAX_STATES; l++) {
        double best = pb::neg_inf();
        uint32_t arg = 7;
        if (l < NS) {
#pragma unroll
            for (int src = 0; src < PB2_MAX_STATES; src++) {
                if ((EDGES >> (8 * l + src)) & 1ull) {
                    const double cand = pb::dadd(pb::dadd(v[src], K.logp[l][src]), e[l]);
                    if (cand > best) { best = cand; arg = src; }
                }
            }
        }
        nv[l] = best;
        bp |= arg << (3 * l);
    }
#pragma unroll
    for (int l = 0; l < PB2_MAX_STATES; l++) v[l] = nv[l];
    return bp;
}

inline uint64_t pack_edges(const HmmMask &K)
{
    uint64_t m = 0;
    for (int l = 0; l < PB2_MAX_STATES; l++) m |= (uint64_t)(K.edge[l] & 0xFFu) << (8 * l);
    return m;
}

__device__ __forceinline__ void viterbi_init(const HmmDev &M, double (&v)[PB2_MAX_STATES],
                                             const double (&e)[PB2_MAX_STATES])
{
#pragma unroll
    for (int s = 0; s < PB2_MAX_STATES; s++) {
        v[s] = pb::neg_inf();
        if (s < M.n_states && M.log_start[s] > pb::neg_inf()) {
            const double cand = pb::dadd(pb::dadd(0.0, M.log_start[s]), e[s]);
            if (cand > v[s]) v[s] = cand;
        }
    }
}

__device__ __forceinline__ int viterbi_end(const HmmDev &M, const double (&v)[PB2_MAX_STATES],
                                           double &best)
{
    int end = 0;
    best = v[0];
#pragma unroll
    for (int s = 1; s < PB2_MAX_STATES; s++)
        if (s < M.n_states && v[s] > best) { best = v[s]; end = s; }
    return end;
}

inline void make_mask(const HmmDev &M, HmmMask &K)
{
    for (int l = 0; l < PB2_MAX_STATES; l++) {
        K.edge[l] = 0;
        for (int s = 0; s < PB2_MAX_STATES; s++) K.logp[l][s] = -INFINITY;
        if (l < M.n_states)
            for (int k = M.in_begin[l]; k < M.in_begin[l + 1]; k++) {
                K.edge[l] |= 1u << M.in_src[k];
                K.logp[l][M.in_src[k]] = M.in_logp[k];
            }
    }
}


}  // namespace pb
