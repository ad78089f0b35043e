// scpp_b200/plugins/plugin_support.hpp — what every model written against the plugin surface shares: Jacobians from the flow map alone,
// and the members a model gets from its GENERATED table (tools/gen_plugin.cpp -> csrc/gen/<model>.inc).
#pragma once

namespace scpp {

// Jacobians of a model that only has a generic-scalar flow map: one dual-number pass per column
template <class Derived, int NX_, int NU_>
struct AutoJacobian {
    struct Lin { double f[NX_]; double A[NX_][NX_]; double B[NX_][NU_]; };
    // zero-order-hold inputs need a strictly feasible placeholder for the unused last input column: not part of the plugin surface
    static constexpr bool ZOH = false;
    SCPP_HD static void zoh_placeholder_input(const double *, double *u) { for (int j = 0; j < NU_; j++) u[j] = 0.; }
    SCPP_HD static void linearize(const double *x, const double *u, const double *par, Lin &L)
    {
        Dual xd[NX_], ud[NU_], fd[NX_];
        for (int j = 0; j < NX_ + NU_; j++) {
            for (int i = 0; i < NX_; i++) xd[i] = Dual(x[i], i == j ? 1. : 0.);
            for (int i = 0; i < NU_; i++) ud[i] = Dual(u[i], NX_ + i == j ? 1. : 0.);
            Derived::template flow_map<Dual>(xd, ud, par, fd);
            for (int i = 0; i < NX_; i++) { if (j < NX_) L.A[i][j] = fd[i].d; else L.B[i][j - NX_] = fd[i].d; L.f[i] = fd[i].v; }
        }
    }
    SCPP_HD static void A_apply(const Lin &L, const double *v, double *o)
    {
        for (int i = 0; i < NX_; i++) { double a = 0; for (int j = 0; j < NX_; j++) a += L.A[i][j] * v[j]; o[i] = a; }
    }
    SCPP_HD static void B_apply(const Lin &L, const double *w, double *o)
    {
        for (int i = 0; i < NX_; i++) { double a = 0; for (int j = 0; j < NU_; j++) a += L.B[i][j] * w[j]; o[i] = a; }
    }
};

// generated tables (scpp_plugin::emit_inc): how a constant slot is filled and which variables are pinned
struct CstRecipe { int kind; int index; double value; };        // kind 0: literal value ; 1: constants[index] * value ; 2 / 3: x_init / x_final[index] * value
struct PinDesc { int idx; int kind; int index; double value; }; // kind 0: literal value ; 1: x_init[index] * value ; 2: x_final[index] * value ; idx < 0 ends the list


} // namespace scpp

// Members generated from csrc/gen/<model>.inc (macro prefix PFX): table sizes, cone_dim / cone_off, crow / row, and
//   setup()          K0: the model's own parameters() (nondimensionalisation, dynamics parameters, constant block), then the recipe
//                    fills the constant slots of the row table (literal | constants[i] * s | x_init[i] * s | x_final[i] * s)
//   constants_from_slots(), initial_guess()   the constant block back out of the slots for the model's initial_trajectory()
//   fixed()          pinned variables of node k from the pin lists (all nodes, then first / last node override)
#define SCPP_PLUGIN_MEMBERS(PFX) \
    static constexpr int NLP = PFX##_NLP, NCONE = PFX##_NCONE, NCR = PFX##_NCR, MAXDIM = PFX##_MAXDIM; \
    static constexpr int NCST = PFX##_NCST; \
    static_assert(NCST + NCONST <= MAX_CST + 8, "constant slots"); \
    SCPP_HD static constexpr int cone_dim(int c) { constexpr int d[NCONE > 0 ? NCONE : 1] = PFX##_CONE_DIMS; return d[c]; } \
    SCPP_HD static constexpr int cone_off(int c) { constexpr int o[NCONE > 0 ? NCONE : 1] = PFX##_CONE_OFFS; return o[c]; } \
    SCPP_HD static constexpr RowDesc crow(int r) { constexpr RowDesc t[NLP + NCR] = PFX##_ROWS; return t[r]; } \
    SCPP_HD static RowDesc row(int r); \
    SCPP_HD static void setup(const ModelParamsHost &P, int nondim, double *xi, double *xf, double *par, double *cst, double *scale) \
    { \
        double constants[NCONST]; \
        parameters(P, nondim, xi, xf, par, constants, scale); \
        constexpr CstRecipe rec[NCST] = PFX##_CST_RECIPE; \
        for (int s = 0; s < MAX_CST; s++) cst[s] = 0.; \
        for (int s = 0; s < NCST; s++) \
            cst[s] = rec[s].kind == 0 ? rec[s].value : rec[s].value * (rec[s].kind == 1 ? constants : (rec[s].kind == 2 ? xi : xf))[rec[s].index]; \
    } \
    SCPP_HD static void constants_from_slots(const double *cst, double *constants) \
    { \
        constexpr CstRecipe rec[NCST] = PFX##_CST_RECIPE; \
        for (int i = 0; i < NCONST; i++) constants[i] = 0.; \
        for (int s = 0; s < NCST; s++) if (rec[s].kind == 1) constants[rec[s].index] = cst[s] / rec[s].value; \
    } \
    SCPP_HD static void initial_guess(const double *xi, const double *xf, const double *cst, int K, int k, double *x, double *u) \
    { \
        double constants[NCONST]; \
        constants_from_slots(cst, constants); \
        initial_trajectory(xi, xf, constants, K, k, x, u); \
    } \
    SCPP_HD static void pin(const PinDesc *l, const double *xi, const double *xf, uint32_t &mask, double *val) \
    { \
        for (int q = 0; l[q].idx >= 0; q++) { \
            mask |= 1u << l[q].idx; \
            val[l[q].idx] = l[q].kind == 0 ? l[q].value : l[q].value * (l[q].kind == 1 ? xi : xf)[l[q].index]; \
        } \
    } \
    SCPP_HD static uint32_t fixed(const ModelParamsHost &, const double *xi, const double *xf, int K, int k, double *val) \
    { \
        const PinDesc all[] = PFX##_PIN_ALL, first[] = PFX##_PIN_FIRST, last[] = PFX##_PIN_LAST; \
        uint32_t mask = 0; \
        for (int i = 0; i < NX + NU; i++) val[i] = 0.; \
        pin(all, xi, xf, mask, val); \
        if (k == 0) pin(first, xi, xf, mask, val); \
        if (k == K - 1) pin(last, xi, xf, mask, val); \
        return mask; \
    }
