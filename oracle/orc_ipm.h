/* oracle/orc_ipm.h — CPU ORACLE (test infrastructure). Internal: conic solve with explicit KKT ordering keys. */
#ifndef ORC_IPM_H
#define ORC_IPM_H
#include "orc.h"
int orc_conic_solve_keys(int n, int p, int m, int l, int ncones, const int *q,
                         const double *c, const double *b, const double *h,
                         int nnzA, const int *Ai, const int *Aj, const double *Av,
                         int nnzG, const int *Gi, const int *Gj, const double *Gv,
                         const double *keys_var, const double *keys_eq,
                         double *x, double *y, double *s, double *z, orc_ipm_info *info);
/* static regularisation of the reduced KKT system (default 1e-13; not thread-safe to change while solves run) */
void orc_set_static_reg(double v);
double orc_get_static_reg(void);
#endif
