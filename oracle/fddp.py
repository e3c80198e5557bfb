"""CPU restatement (numpy) of SolverFDDPTpl::backwardPass, solvers/fddp/solver-fddp.hxx:204-277 -- TEST
INFRASTRUCTURE (same rule as oracle/gar_oracle.hpp; parity unpinned: the reference cannot run here).
Statement by statement: terminal value (:212-222), Q-function assembly (:240-248), gains through an
LLT of Quu (:252-262), Quuks (:264), value function with the selfadjoint-lower symmetrisation, the
regularisation of its diagonal and the defect term Vx += Vxx fs[i] (:266-276)."""
import numpy as np


def backward_pass(Jx, Ju, fs, Lxx, Lxu, Luu, Lx, Lu, Lxx_N, Lx_N, preg):
    """Per-stage lists (length N) of Jx [nx,nx], Ju [nx,nu], Lxx, Lxu [nx,nu], Luu, Lx, Lu; fs: N+1 defects.
    Returns dict of lists: K (kkt_fb), k (kkt_ff), Vxx, Vx (N+1 entries), Quuks (N)."""
    N = len(Jx)
    nx = Lxx_N.shape[0]
    Vxx = [None] * (N + 1)
    Vx = [None] * (N + 1)
    K, k, Quuks = [None] * N, [None] * N, [None] * N
    V = Lxx_N.copy()
    V[np.diag_indices(nx)] += preg                                   # :217
    Vxx[N] = V
    Vx[N] = Lx_N + V @ fs[N]                                          # :216, :219-220
    for i in range(N - 1, -1, -1):
        J = np.hstack([Jx[i], Ju[i]])                                 # J_x_u (:236)
        nu = Ju[i].shape[1]
        grad = np.concatenate([Lx[i], Lu[i]]) + J.T @ Vx[i + 1]       # :239-240
        hess = np.block([[Lxx[i], Lxu[i]], [Lxu[i].T, Luu[i]]]) + (J.T @ Vxx[i + 1]) @ J   # :243-245
        Qxx, Qxu, Quu = hess[:nx, :nx], hess[:nx, nx:], hess[nx:, nx:].copy()
        Quu[np.diag_indices(nu)] += preg                              # :246
        Qx, Qu = grad[:nx], grad[nx:]
        L = np.linalg.cholesky(Quu)                                   # :259-260 (LLT)
        sol = np.linalg.solve(L.T, np.linalg.solve(L, np.column_stack([-Qu, -Qxu.T])))   # :255-261
        kff, kfb = sol[:, 0], sol[:, 1:]
        k[i], K[i] = kff, kfb
        Quuks[i] = Quu @ kff                                          # :264
        vx = Qx + kfb.T @ Qu                                          # :268-269
        v = Qxx + Qxu @ kfb                                           # :270-271
        v = np.tril(v) + np.tril(v, -1).T                             # :272 selfadjointView<Lower>
        v[np.diag_indices(nx)] += preg                                # :273
        Vxx[i] = v
        Vx[i] = vx + v @ fs[i]                                        # :274-276
    return dict(K=K, k=k, Vxx=Vxx, Vx=Vx, Quuks=Quuks)
