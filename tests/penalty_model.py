"""numpy restatement (TEST INFRASTRUCTURE) of the reference's Mehrotra loop for EqualityHandling::PenaltyFunction /
PenaltyFunctionWithExtraDual, which the reference derives symbolically but cannot run: its evaluator has no rule for the
scalar diagonal block -mu I (Evaluation.cpp:53-60).

System (printed by the reference's own symbolic layer for Settings{inequalities = None, variable_bounds = Both,
equalities = true, equality_handling = PenaltyFunction*}; get_newton_system / get_augmented_system, both handlings give
the same rows):

    [ Q + Y^-1 L_y + Z^-1 L_z   C^T ] [dx  ]   [ Z^-1 (r_z - L_z r_lz) - r_x - Y^-1 (r_y - L_y r_ly) ]
    [ C                        -mu  ] [dlam] = [ -(C x - d - mu lam) ]
    r_x = c + lam_z + Q x + C^T lam - lam_y,  r_ly = l + y - x,  r_lz = x + z - u,  r_y = Y lam_y - mu e,  r_z = Z lam_z - mu e
    dlam_y = -Y^-1 L_y (dx - r_ly + L_y^-1 r_y),  dy = -L_y^-1 (Y dlam_y + r_y)   (likewise z with the opposite sign of dx)

mu in the loop (Optimizer.cpp:126-218): the matrix is evaluated with the value the previous iteration left in the
environment (sigma mu; 1 before the first one, EnvironmentBuilder.cpp:48); the predictor's residuals, `res` and `gap`
with mu = 0; the corrector's with the new sigma mu (complementarity rows additionally get dv_aff * dlam_aff).  One step
length for all variables, capped at 1, x clamped to its bounds directly (no g / h slacks: Optimizer.cpp:296-317),
times 0.995."""
import numpy as np


def _ratio(v, dv, a):
    m = dv < 0
    return min(a, float(np.min(-v[m] / dv[m]))) if np.any(m) else a


def solve(Q, c, C, d, l, u, max_iter=100, tol=1e-8):
    n, me = Q.shape[0], C.shape[0]
    x = 0.5 * (l + u)
    lam, ly, lz, y, z = np.ones(me), np.ones(n), np.ones(n), np.ones(n), np.ones(n)
    mu_env = 1.0
    tr = dict(f=[], res=[], mu=[], step_aff=[], step_cor=[], alpha_aff=[], sigma=[], alpha=[])
    it = 0
    for it in range(max_iter + 1):
        qx = Q @ x
        f = float(np.dot(0.5 * x, qx) + np.dot(c, x))
        r_x = c + lz + qx + C.T @ lam - ly
        r_lam0 = C @ x - d
        r_ly, r_lz = l + y - x, x + z - u
        r_y, r_z = y * ly, z * lz
        res = float(np.sqrt(sum(np.dot(r, r) for r in (r_x, r_lam0, r_ly, r_lz, r_y, r_z))))
        mu = float((np.sum(np.abs(r_y)) + np.sum(np.abs(r_z))) / (2 * n))
        tr["f"].append(f); tr["res"].append(res); tr["mu"].append(mu)
        if (res < tol and mu < tol) or it == max_iter:
            break
        K = np.block([[Q + np.diag(ly / y) + np.diag(lz / z), C.T], [C, -mu_env * np.eye(me)]])

        def direction(r_y, r_z, r_lam):
            b0 = (r_z - lz * r_lz) / z - r_x - (r_y - ly * r_ly) / y
            s = np.linalg.solve(K, np.concatenate([b0, -r_lam]))
            dx, dlam = s[:n], s[n:]
            dly = -(ly / y) * (dx - r_ly + r_y / ly)
            dy = -(y * dly + r_y) / ly
            dlz = -(lz / z) * (-r_lz - dx + r_z / lz)
            dz = -(z * dlz + r_z) / lz
            a = 1.0
            for v, dv in ((y, dy), (z, dz), (ly, dly), (lz, dlz)):
                a = _ratio(v, dv, a)
            m = dx < 0
            if np.any(m): a = min(a, float(np.min((l[m] - x[m]) / dx[m])))
            m = dx > 0
            if np.any(m): a = min(a, float(np.min((u[m] - x[m]) / dx[m])))
            return s, dx, dlam, dly, dlz, dy, dz, a

        sa, dxa, dlama, dlya, dlza, dya, dza, a_aff = direction(r_y, r_z, r_lam0)
        mu_aff = float((np.sum(np.abs((y + a_aff * dya) * (ly + a_aff * dlya))) +
                        np.sum(np.abs((z + a_aff * dza) * (lz + a_aff * dlza)))) / (2 * n))
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        mu_c = mu * sigma
        sc, dx, dlam, dly, dlz, dy, dz, a = direction(y * ly - mu_c + dya * dlya, z * lz - mu_c + dza * dlza,
                                                      r_lam0 - mu_c * lam)
        tr["step_aff"].append(sa); tr["step_cor"].append(sc)
        tr["alpha_aff"].append(a_aff); tr["sigma"].append(sigma); tr["alpha"].append(a)
        st = 0.995 * a
        x, lam, ly, lz, y, z = x + st * dx, lam + st * dlam, ly + st * dly, lz + st * dlz, y + st * dy, z + st * dz
        mu_env = mu_c
    tr["iterations"] = it
    tr["converged"] = bool(tr["res"][-1] < tol and tr["mu"][-1] < tol)
    tr["x"], tr["lam"] = x, lam
    return tr
