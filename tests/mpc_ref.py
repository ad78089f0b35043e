"""TEST-ONLY: the MPC sub-problem of the reference restated for the oracle's generic conic solver (orc_conic_solve), plus the condensed
form the engine solves, both built here in numpy from the same inputs.

Follows, line by line: exactLinearDiscretization (scpp_core/src/discretization.cpp:9-40, Eigen's matrix exponential = scipy.linalg.expm),
buildMPCProblem (scpp_core/src/MPCProblem.cpp:6-87), Rocket2d::addApplicationConstraints (scpp_models/src/rocket2d.cpp:46-84) with
constrain_initial_final = false, Rocket2d::getOperatingPoint (:40-44, the intended hover point: the reference's comma initialiser writes three
coefficients into a 2-vector)."""
import numpy as np
import scipy.linalg as sla


def r2d_f(x, u, m, J, g, rT):
    TBx, TBy = -np.sin(u[0]) * u[1], np.cos(u[0]) * u[1]
    ce, se = np.cos(x[4]), np.sin(x[4])
    return np.array([x[2], x[3], (ce * TBx - se * TBy) / m + g[0], (se * TBx + ce * TBy) / m + g[1], x[5], (rT[0] * TBy - rT[1] * TBx) / J])


def discretize(p, ts):
    """A, B, z of exactLinearDiscretization at the operating point (central differences of the flow map stand in for CppAD: the dynamics are
    smooth and the step is chosen for 1e-10 accuracy)"""
    m, J, g, rT = p.m, p.J_B, np.array(p.g_I), np.array(p.r_T_B)
    xe = np.zeros(6); ue = np.array([0.0, -g[1] * m])
    f = lambda x, u: r2d_f(x, u, m, J, g, rT)
    Ac = np.zeros((6, 6)); Bc = np.zeros((6, 2))
    for j in range(6):
        h = 1e-6; e = np.zeros(6); e[j] = h
        Ac[:, j] = (f(xe + e, ue) - f(xe - e, ue)) / (2 * h)
    for j in range(2):
        h = 1e-6 * max(1.0, abs(ue[j])); e = np.zeros(2); e[j] = h
        Bc[:, j] = (f(xe, ue + e) - f(xe, ue - e)) / (2 * h)
    E = np.zeros((8, 8)); E[:6, :6] = Ac; E[:6, 6:] = Bc
    X = sla.expm(E * ts)
    A, B = X[:6, :6], X[:6, 6:]
    E = np.zeros((7, 7)); E[:6, :6] = Ac; E[:6, 6] = f(xe, ue) - Ac @ xe - Bc @ ue
    z = sla.expm(E * ts)[:6, 6]
    return A, B, z


def full_socp(p, K, A, B, z, x_init, x_final, w_term, w_in, state_rows_from=0):
    """ECOS standard form  min c'x  s.t.  Aeq x = b,  h - G x in K  of buildMPCProblem + addApplicationConstraints.
    Variables: X (column k at 6k..6k+5), U (column k at 6K + 2k), error_cost, input_cost."""
    nx, nu = 6, 2
    n = nx * K + nu * (K - 1) + 2
    iX = lambda k, i: nx * k + i
    iU = lambda k, i: nx * K + nu * k + i
    ie, ic = n - 2, n - 1
    c = np.zeros(n); c[ie] = 1; c[ic] = 1
    Ae, be = [], []
    for i in range(nx):                                   # X.col(0) == x_init   (MPCProblem.cpp:27-30)
        r = np.zeros(n); r[iX(0, i)] = 1; Ae.append(r); be.append(x_init[i])
    for k in range(K - 1):                                # A x_k + B u_k + z == x_{k+1}   (:32-55)
        for i in range(nx):
            r = np.zeros(n)
            for j in range(nx): r[iX(k, j)] += A[i, j]
            for j in range(nu): r[iU(k, j)] += B[i, j]
            r[iX(k + 1, i)] -= 1
            Ae.append(r); be.append(-z[i])
    G, h, q = [], [], []
    def lp(row, hv): G.append(row); h.append(hv)
    tg = np.tan(p.gamma_gs)
    for k in range(state_rows_from, K):                   # boxes on eta and omega (rocket2d.cpp:66-72)
        for i, lim in ((4, p.theta_max), (5, p.w_B_max)):
            r = np.zeros(n); r[iX(k, i)] = 1; lp(r, lim)
            r = np.zeros(n); r[iX(k, i)] = -1; lp(r, lim)
    for k in range(K - 1):                                # gimbal and thrust ranges (:76-82)
        r = np.zeros(n); r[iU(k, 0)] = 1; lp(r, p.gimbal_max)
        r = np.zeros(n); r[iU(k, 0)] = -1; lp(r, p.gimbal_max)
        r = np.zeros(n); r[iU(k, 1)] = -1; lp(r, -p.T_min)
        r = np.zeros(n); r[iU(k, 1)] = 1; lp(r, p.T_max)
    l = len(G)
    for k in range(state_rows_from, K):                   # glide slope |r_x| <= tan(gamma) r_y   (:63-64)
        r = np.zeros(n); r[iX(k, 1)] = -tg; G.append(r); h.append(0.)
        r = np.zeros(n); r[iX(k, 0)] = -1; G.append(r); h.append(0.)
        q.append(2)
    r = np.zeros(n); r[ie] = -1; G.append(r); h.append(0.)          # |W_T (X_{K-1} - x_final)| <= error_cost   (:61-73)
    for i in range(nx):
        r = np.zeros(n); r[iX(K - 1, i)] = -w_term[i]; G.append(r); h.append(-w_term[i] * x_final[i])
    q.append(1 + nx)
    r = np.zeros(n); r[ic] = -1; G.append(r); h.append(0.)          # |W_u U| <= input_cost   (:79-86)
    for k in range(K - 1):
        for j in range(nu):
            r = np.zeros(n); r[iU(k, j)] = -w_in[j]; G.append(r); h.append(0.)
    q.append(1 + nu * (K - 1))
    return dict(c=c, A=np.array(Ae), b=np.array(be), G=np.array(G), h=np.array(h), l=l, q=q, iX=iX, iU=iU, n=n)


def solve_with_oracle(O, P):
    import scipy.sparse as sp
    A = sp.coo_matrix(P["A"]); G = sp.coo_matrix(P["G"])
    r = O.conic_solve(P["c"], P["b"], P["h"], P["l"], P["q"], (A.row, A.col, A.data), (G.row, G.col, G.data))
    return r


def condensed(p, K, A, B, z, x_init, x_final, w_term, w_in):
    """the engine's form: y = (U, error_cost, input_cost), x_k = Phi_k x0 + S_k U + zh_k; state rows at nodes 1..K-1.  Returns nv, nl, cdim, G, c, h"""
    nx, nu = 6, 2
    nuu = nu * (K - 1); nv = nuu + 2
    Phi = [np.eye(nx)]; S = [np.zeros((nx, nuu))]; zh = [np.zeros(nx)]
    for k in range(K - 1):
        Sk = A @ S[k]; Sk[:, nu * k:nu * k + nu] += B
        Phi.append(A @ Phi[k]); S.append(Sk); zh.append(A @ zh[k] + z)
    xk = lambda k: Phi[k] @ x_init + zh[k]
    G, h = [], []
    tg = np.tan(p.gamma_gs)
    for k in range(1, K):
        for i, lim in ((4, p.theta_max), (5, p.w_B_max)):
            r = np.zeros(nv); r[:nuu] = S[k][i]; G.append(r); h.append(lim - xk(k)[i])
            r = np.zeros(nv); r[:nuu] = -S[k][i]; G.append(r); h.append(lim + xk(k)[i])
    for k in range(K - 1):
        r = np.zeros(nv); r[nu * k] = 1; G.append(r); h.append(p.gimbal_max)
        r = np.zeros(nv); r[nu * k] = -1; G.append(r); h.append(p.gimbal_max)
        r = np.zeros(nv); r[nu * k + 1] = -1; G.append(r); h.append(-p.T_min)
        r = np.zeros(nv); r[nu * k + 1] = 1; G.append(r); h.append(p.T_max)
    nl = len(G); cdim = []
    for k in range(1, K):
        r = np.zeros(nv); r[:nuu] = -tg * S[k][1]; G.append(r); h.append(tg * xk(k)[1])
        r = np.zeros(nv); r[:nuu] = -S[k][0]; G.append(r); h.append(xk(k)[0])
        cdim.append(2)
    r = np.zeros(nv); r[nuu] = -1; G.append(r); h.append(0.)
    for i in range(nx):
        r = np.zeros(nv); r[:nuu] = -w_term[i] * S[K - 1][i]; G.append(r); h.append(w_term[i] * (xk(K - 1)[i] - x_final[i]))
    cdim.append(1 + nx)
    r = np.zeros(nv); r[nuu + 1] = -1; G.append(r); h.append(0.)
    for k in range(K - 1):
        for j in range(nu):
            r = np.zeros(nv); r[nu * k + j] = -w_in[j]; G.append(r); h.append(0.)
    cdim.append(1 + nuu)
    c = np.zeros(nv); c[nuu] = 1; c[nuu + 1] = 1
    return nv, nl, cdim, np.array(G), c, np.array(h), (Phi, S, zh)
