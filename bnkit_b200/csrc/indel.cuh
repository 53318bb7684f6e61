// indel.cuh -- types shared by indel_tables.cu (table construction, -fmad=false) and indel.cu (the sweep and the
// bnk_indel_* entry points): gap-augmented peeling of asr.IndelPeeler / bn.ctmc.GapSubstModel.
#pragma once
#include "device_common.cuh"

namespace bnk {

struct GapModelDev {   // device pointers; K1 = residues + gap
    const double *evec, *ievec, *eval;  // K1*K1, K1*K1, K1
    const double *F;                    // K1 stationary frequencies incl. the gap entry
    double mu, lambda;
    int K1;
};

// per-branch tables for one (mu, lambda, rate); K = residues
struct IndelTables {
    double *A;    // [node][x][p] = P_eps(t_node)[p][x] * (1 - ksi(t_node))     residue block only; x = child residue, p = parent residue
    double *del;  // [node]       = (1 - ksi(t_node)) * gamma(t_node)
    double *ins;  // [node][q]    = ksi(t_node) * F[q]           read when `node` is the PARENT (IndelPeeler.java:392-400)
};

void launch_indel_tables(int K1, const GapModelDev &gm, const double *dist, const int32_t *parent, double rate, int n_nodes,
                         const IndelTables &tb, cudaStream_t st);
void launch_indel_probs(int K1, const GapModelDev &gm, double t, double *P, cudaStream_t st);

}  // namespace bnk
