// subst_model.cpp -- host-side substitution-model setup (once per model; 20x20).
//
// Mirrors bn.ctmc.SubstModel's constructor chain (src/bn/ctmc/SubstModel.java:80-110 R from
// S*F, :192-202 makeValid, :209-219 normalize) and bn.math.Matrix.Exp (Matrix.java:42-60).
// Everything downstream (P(t), log P(t)) is built on the device from the eigensystem kept
// here.  Compiled with -ffp-contract=off so each operation rounds separately (as on a JVM).
#include "internal.h"

#include <cmath>
#include <cstring>
#include <new>

#define BNK_EIG_MAXK (BNK_MAXK + 1)  // + the gap state of bn.ctmc.GapSubstModel

namespace bnk {
namespace eig {
#include "subst_eigen.inc"
#undef A_
#undef H_
#undef Z_
}  // namespace eig

namespace data {
#include "model_data.inc"
}

// Matrix.java:42-60: copy, elmhes, eltran, hqr2(dim, 1, dim), luinverse
static int eigensystem(int n, const double *A, double *evec, double *wr, double *wi, double *ievec)
{
    double work[BNK_EIG_MAXK * BNK_EIG_MAXK];
    int ordr[BNK_EIG_MAXK];
    std::memcpy(work, A, sizeof(double) * n * n);
    std::memset(evec, 0, sizeof(double) * n * n);
    eig::elmhes(work, ordr, n);
    eig::eltran(work, evec, ordr, n);
    if (eig::hqr2(n, 1, n, work, evec, wr, wi) != 0) return -1;
    if (eig::luinverse(evec, ievec, n) != 0) return -2;
    return 0;
}

static void set_prior(bnk_model *m)
{
    // new EnumDistrib(domain, F): copy then divide by the sum (SubstNode.java:109,
    // EnumDistrib.java:66-78, :363-371)
    double sum = 0.0;
    for (int i = 0; i < m->K; i++) sum += m->F[i];
    for (int i = 0; i < m->K; i++) m->prior[i] = m->F[i] / sum;
}

int model_from_rates(int K, const double *F, const double *M, bool symmetric, bool normalise,
                     bnk_model **out)
{
    if (!out || !F || !M) return fail(BNK_ERR_ARG, "model: null argument");
    if (K < 2 || K > BNK_MAXK) return fail(BNK_ERR_ARG, "model: n_states must be in 2..%d", BNK_MAXK);
    bnk_model *m = new (std::nothrow) bnk_model();
    if (!m) return fail(BNK_ERR_NOMEM, "model: out of host memory");
    m->K = K;
    for (int i = 0; i < K; i++) m->F[i] = F[i];
    double *R = m->R;
    for (int i = 0; i < K; i++) {
        if (symmetric) {
            for (int j = i + 1; j < K; j++) {
                const double s = M[i * K + j];
                R[i * K + j] = s * F[j];
                R[j * K + i] = s * F[i];
            }
        } else {
            for (int j = 0; j < K; j++)
                if (i != j) R[i * K + j] = M[i * K + j];
        }
    }
    for (int i = 0; i < K; i++) {  // makeValid
        double sum = 0.0;
        for (int j = 0; j < K; j++)
            if (i != j) sum += R[i * K + j];
        R[i * K + i] = -sum;
    }
    if (normalise) {
        double sum = 0.0;
        for (int i = 0; i < K; i++) sum += -R[i * K + i] * F[i];
        for (int i = 0; i < K * K; i++) R[i] = R[i] / sum;
    }
    double wi[BNK_MAXK];
    const int rc = eigensystem(K, R, m->evec, m->eval, wi, m->ievec);
    if (rc != 0) {
        delete m;
        return fail(BNK_ERR_MODEL, rc == -1 ? "model: eigenvalues not converged" : "model: singular eigenvector matrix");
    }
    set_prior(m);
    *out = m;
    return BNK_OK;
}

int model_by_name(const char *name, const double *F_override, bnk_model **out)
{
    if (!name || !out) return fail(BNK_ERR_ARG, "model: null argument");
    const double *F = nullptr, *S = nullptr;
    if (!std::strcmp(name, "LG")) { F = data::BNK_LG_F; S = data::BNK_LG_S; }
    else if (!std::strcmp(name, "JTT")) { F = data::BNK_JTT_F; S = data::BNK_JTT_S; }
    else if (!std::strcmp(name, "WAG")) { F = data::BNK_WAG_F; S = data::BNK_WAG_S; }
    else if (!std::strcmp(name, "Dayhoff")) { F = data::BNK_DAYHOFF_F; S = data::BNK_DAYHOFF_S; }
    int rc;
    if (F) {
        double M[400] = {0.0};
        int k = 0;
        for (int i = 0; i < 20; i++)
            for (int j = i + 1; j < 20; j++) M[i * 20 + j] = S[k++];
        rc = model_from_rates(20, F_override ? F_override : F, M, true, true, out);
    } else if (!std::strcmp(name, "Yang")) {  // Yang.java:32-50
        rc = model_from_rates(4, F_override ? F_override : data::BNK_YANG_F, data::BNK_YANG_Q, false, true, out);
    } else if (!std::strcmp(name, "JC")) {    // JC.java:32-57: gamma = 1, not normalised
        double Fj[4], Qj[16];
        for (int r = 0; r < 4; r++) {
            Fj[r] = 1.0 / 4;
            for (int c = 0; c < 4; c++) Qj[r * 4 + c] = (r != c) ? 1.0 / 4 : 0.0;
        }
        rc = model_from_rates(4, Fj, Qj, false, false, out);
    } else {
        return fail(BNK_ERR_MODEL, "model: unknown name '%s' (LG, JTT, WAG, Dayhoff, JC, Yang)", name);
    }
    if (rc == BNK_OK) std::snprintf((*out)->name, sizeof((*out)->name), "%s", name);
    return rc;
}

// The base model IndelPeeler.optimiseMuLambda builds for a model name (src/asr/IndelPeeler.java:474-503): always
// new GapSubstModel(F, IRM, alpha, mu, lambda) = symmetric (upper triangle x F) AND normalised -- for JC and Yang
// that is NOT the model SubstModel.createModel(name) gives (JC: unnormalised; Yang: its Q taken as a full matrix).
int indel_base_model_by_name(const char *name, bnk_model **out)
{
    if (!name || !out) return fail(BNK_ERR_ARG, "model: null argument");
    int rc;
    if (!std::strcmp(name, "JC")) {
        double Fj[4], Qj[16];
        for (int r = 0; r < 4; r++) {
            Fj[r] = 1.0 / 4;
            for (int c = 0; c < 4; c++) Qj[r * 4 + c] = (r != c) ? 1.0 / 4 : 0.0;
        }
        rc = model_from_rates(4, Fj, Qj, true, true, out);
    } else if (!std::strcmp(name, "Yang")) {
        rc = model_from_rates(4, data::BNK_YANG_F, data::BNK_YANG_Q, true, true, out);
    } else if (!std::strcmp(name, "LG") || !std::strcmp(name, "JTT") || !std::strcmp(name, "WAG") || !std::strcmp(name, "Dayhoff")) {
        return model_by_name(name, nullptr, out);   // already symmetric + normalised
    } else {
        return fail(BNK_ERR_MODEL, "%s gap model not implemented yet", name);   // IllegalArgumentException, :502
    }
    if (rc == BNK_OK) std::snprintf((*out)->name, sizeof((*out)->name), "%s", name);
    return rc;
}

// new GapSubstModel(F, IRM, alphabet, mu, lambda) on top of a constructed base model
// (src/bn/ctmc/GapSubstModel.java:26-44): R_eps = [R - mu I | mu ; lambda * pi | -lambda] (constructIndelR, :61-86),
// Exp(R_eps), stationary frequencies with the gap entry mu / (mu + lambda) (addGapToStationaryFreqs, :104-131)
int gap_model_from_base(const bnk_model *base, double mu, double lambda, GapModelHost *g)
{
    if (!base || !g) return fail(BNK_ERR_ARG, "gap model: null argument");
    if (!base->has_R) return fail(BNK_ERR_MODEL, "gap model: the base model was adopted from an eigensystem and has no rate matrix");
    if (!(mu + lambda >= 0)) return fail(BNK_ERR_ARG, "gap model: mu + lambda must be >= 0");
    const int K = base->K, K1 = K + 1;
    g->K1 = K1; g->mu = mu; g->lambda = lambda;
    std::memset(g->R, 0, sizeof(g->R));
    for (int i = 0; i < K; i++) {
        g->R[i * K1 + K] = mu;
        for (int j = 0; j < K; j++) {
            g->R[i * K1 + j] = base->R[i * K + j];
            if (i == j) g->R[i * K1 + j] -= mu;
        }
    }
    for (int j = 0; j < K; j++) g->R[K * K1 + j] = base->F[j] * lambda;
    g->R[K * K1 + K] = -lambda;
    double wi[BNK_EIG_MAXK];
    const int rc = eigensystem(K1, g->R, g->evec, g->eval, wi, g->ievec);
    if (rc != 0) return fail(BNK_ERR_MODEL, rc == -1 ? "gap model: eigenvalues not converged" : "gap model: singular eigenvector matrix");
    if (mu + lambda > 0) {
        const double gapFreq = mu / (mu + lambda);
        for (int i = 0; i < K; i++) g->F[i] = (1 - gapFreq) * base->F[i];
        g->F[K] = gapFreq;
    } else {
        for (int i = 0; i < K; i++) g->F[i] = base->F[i];
        g->F[K] = 0.0;
    }
    return BNK_OK;
}

int model_from_eigen(int K, const double *F, const double *evec, const double *ievec,
                     const double *eval, bnk_model **out)
{
    if (!out || !F || !evec || !ievec || !eval) return fail(BNK_ERR_ARG, "model: null argument");
    if (K < 2 || K > BNK_MAXK) return fail(BNK_ERR_ARG, "model: n_states must be in 2..%d", BNK_MAXK);
    bnk_model *m = new (std::nothrow) bnk_model();
    if (!m) return fail(BNK_ERR_NOMEM, "model: out of host memory");
    m->K = K;
    std::memcpy(m->F, F, sizeof(double) * K);
    std::memcpy(m->evec, evec, sizeof(double) * K * K);
    std::memcpy(m->ievec, ievec, sizeof(double) * K * K);
    std::memcpy(m->eval, eval, sizeof(double) * K);
    m->has_R = false;
    set_prior(m);
    *out = m;
    return BNK_OK;
}

}  // namespace bnk
