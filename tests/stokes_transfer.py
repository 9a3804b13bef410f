"""Stokes transfer along a ray in a chosen floating-point type (TEST INFRASTRUCTURE).

numpy restatement of the reference's per-sample coupling of the Stokes vector to the plasma without rotation splitting
(src/radiation_integrator/polarized.cpp:569-790: the five analytic cases, the admissibility clamp) applied to the
per-sample inputs the CUDA pipeline leaves in its scratch -- transport matrix M, affine step, the eight synchrotron
coefficients of one frequency (Context.polarized_scratch).  Run in float64 and in numpy.longdouble (x87 80-bit, 64-bit
mantissa) on the SAME inputs, the difference between the two results is the round-off that the closed forms amplify:
where it is large, two correct double-precision implementations of these formulas legitimately disagree by as much."""
import numpy as np

F_M, F_DLAM, F_COEF = 0, 10, 18


def couple(s, j, al, rho, dl, xp):
    """One sample.  s, j, al, rho: (R, 4) (Stokes I, Q, U, V order; the U entries of j, al, rho are zero), dl: (R,)."""
    one = xp(1.0)
    al0 = al[:, 0]
    dt = al0 * dl
    thin = dt <= 100.0
    alpha_sq = al[:, 1] ** 2 + al[:, 3] ** 2
    alpha_p = np.sqrt(alpha_sq)
    rho_sq = rho[:, 1] ** 2 + rho[:, 3] ** 2
    rho_p = np.sqrt(rho_sq)
    out = np.zeros_like(s)
    with np.errstate(all='ignore'):
        # A: no absorptivity, no rotativity
        case_a = (al0 == 0) & (rho_p == 0)
        out_a = s + j * dl[:, None]
        # B: unpolarized absorptivity only
        case_b = ~case_a & (alpha_p == 0) & (rho_p == 0)
        out_b = np.where(thin[:, None], np.exp(-dt)[:, None] * (s + j / al0[:, None] * np.expm1(dt)[:, None]), j / al0[:, None])
        # C: rotativity without absorptivity (I A2-A5)
        case_c = ~case_a & ~case_b & (al0 == 0)
        cr, sr = np.cos(rho_p * dl), np.sin(rho_p * dl)
        ssq = np.sin(rho_p * dl / 2) ** 2
        rho_ss = rho[:, 1] * s[:, 1] + rho[:, 3] * s[:, 3]
        out_c = np.stack([s[:, 0],
                          s[:, 1] * cr + 2 * rho[:, 1] * rho_ss / rho_sq * ssq - rho[:, 3] * s[:, 2] / rho_p * sr,
                          s[:, 2] * cr + (rho[:, 3] * s[:, 1] - rho[:, 1] * s[:, 3]) / rho_p * sr,
                          s[:, 3] * cr + 2 * rho[:, 3] * rho_ss / rho_sq * ssq + rho[:, 1] * s[:, 2] / rho_p * sr], axis=1) + j * dl[:, None]
        # D: polarized absorptivity without rotativity (I A14-A17)
        case_d = ~case_a & ~case_b & ~case_c & (rho_p == 0)
        xq = alpha_p * dl
        e_i, e_p = np.exp(-dt), np.exp(-xq)
        sh, ch = np.sinh(xq), np.cosh(xq)
        chm1 = 0.5 * (np.expm1(xq) + e_p - one)
        a_ss = al[:, 1] * s[:, 1] + al[:, 3] * s[:, 3]
        a_j = al[:, 1] * j[:, 1] + al[:, 3] * j[:, 3]
        fac = one / (al0 * al0 - alpha_sq)
        d0 = (s[:, 0] * ch - a_ss / alpha_p * sh) * e_i + a_j * fac * (-one + (al0 * sh + alpha_p * ch) / alpha_p * e_p) \
            + al0 * j[:, 0] * fac * (one - (al0 * ch + alpha_p * sh) / al0 * e_p)
        cols = [d0]
        for a in (1, 2, 3):
            t1 = (s[:, a] + al[:, a] * a_ss / alpha_sq * chm1 - s[:, 0] * al[:, a] / alpha_p * sh) * e_i
            t2 = j[:, a] * (one - e_i) / al0
            t3 = a_j * al[:, a] / al0 * fac * (one - (one - al0 * al0 / alpha_sq - al0 / alpha_sq * (al0 * ch + alpha_p * sh)) * e_i)
            t4 = j[:, 0] * al[:, a] / alpha_p * fac * (-alpha_p + (alpha_p * ch + al0 * sh) * e_i)
            cols.append(t1 + t2 + t3 + t4)
        thin_d = np.stack(cols, axis=1)
        k0 = (al0 * j[:, 0] - a_j) / (al0 * al0 - alpha_sq)
        thick_d = np.stack([k0] + [(j[:, a] - al[:, a] * k0) / al0 for a in (1, 2, 3)], axis=1)
        out_d = np.where(thin[:, None], thin_d, thick_d)
        # E: absorptivity and rotativity (polarized.cpp:656-779, with the entries the reference leaves unset at zero)
        a_rho = al[:, 1] * rho[:, 1] + al[:, 3] * rho[:, 3]
        dd = alpha_sq - rho_sq
        lam_a = np.sqrt(dd * dd / 4 + a_rho * a_rho)
        lam_b = dd / 2
        l1, l2 = np.sqrt(lam_a + lam_b), np.sqrt(lam_a - lam_b)
        theta = l1 * l1 + l2 * l2
        sg = np.where(a_rho >= 0, one, -one)
        R = len(dl)
        m1 = np.zeros((R, 4, 4), dtype=s.dtype)
        m2, m3, m4 = m1.copy(), m1.copy(), m1.copy()
        for a in range(4):
            m1[:, a, a] = one
        m2[:, 0, 1] = l2 * al[:, 1] - sg * l1 * rho[:, 1]
        m2[:, 0, 3] = l2 * al[:, 3] - sg * l1 * rho[:, 3]
        m2[:, 1, 2] = sg * l1 * al[:, 1] + l2 * rho[:, 1]
        m2[:, 1, 0], m2[:, 3, 0], m2[:, 2, 1] = m2[:, 0, 1], m2[:, 0, 3], -m2[:, 1, 2]
        m3[:, 0, 1] = l1 * al[:, 1] + sg * l2 * rho[:, 1]
        m3[:, 0, 3] = l1 * al[:, 3] + sg * l2 * rho[:, 3]
        m3[:, 1, 2] = -(sg * l2 * al[:, 1] - l1 * rho[:, 1])
        m3[:, 1, 0], m3[:, 3, 0], m3[:, 2, 1] = m3[:, 0, 1], m3[:, 0, 3], -m3[:, 1, 2]
        half = (alpha_sq + rho_sq) / 2
        m4[:, 0, 0], m4[:, 2, 2] = half, -half
        m4[:, 1, 1] = al[:, 1] ** 2 + rho[:, 1] ** 2 - half
        m4[:, 3, 3] = al[:, 3] ** 2 + rho[:, 3] ** 2 - half
        m4[:, 0, 2] = al[:, 1] * rho[:, 3] - al[:, 3] * rho[:, 1]
        m4[:, 1, 3] = al[:, 3] * al[:, 1] + rho[:, 3] * rho[:, 1]
        m4[:, 2, 0], m4[:, 3, 1] = -m4[:, 0, 2], m4[:, 1, 3]
        m2 *= (one / theta)[:, None, None]
        m3 *= (one / theta)[:, None, None]
        m4 *= (2 / theta)[:, None, None]
        ex = np.exp(-dt)[:, None, None]
        sn, cs = np.sin(l2 * dl)[:, None, None], np.cos(l2 * dl)[:, None, None]
        snh, csh = np.sinh(l1 * dl)[:, None, None], np.cosh(l1 * dl)[:, None, None]
        oo = ex * (0.5 * (m1 + m4) * csh + 0.5 * (m1 - m4) * cs - m2 * sn - m3 * snh)
        f1 = (one / (al0 * al0 - l1 * l1))[:, None, None]
        f2 = (one / (al0 * al0 + l2 * l2))[:, None, None]
        a0, L1, L2 = al0[:, None, None], l1[:, None, None], l2[:, None, None]
        cosh_t = -L1 * f1 * m3 + 0.5 * a0 * f1 * (m1 + m4)
        cos_t = -L2 * f2 * m2 + 0.5 * a0 * f2 * (m1 - m4)
        sin_t = -a0 * f2 * m2 - 0.5 * L2 * f2 * (m1 - m4)
        sinh_t = -a0 * f1 * m3 + 0.5 * L1 * f1 * (m1 + m4)
        pp_thick = cosh_t + cos_t
        pp_thin = pp_thick - ex * (cosh_t * csh + cos_t * cs + sin_t * sn + sinh_t * snh)
        e_thin = np.einsum('rab,rb->ra', pp_thin, j) + np.einsum('rab,rb->ra', oo, s)
        e_thick = np.einsum('rab,rb->ra', pp_thick, j)
        out_e = np.where(thin[:, None], e_thin, e_thick)
    out = np.where(case_a[:, None], out_a, np.where(case_b[:, None], out_b, np.where(case_c[:, None], out_c,
                   np.where(case_d[:, None], out_d, out_e))))
    # admissibility (polarized.cpp:782-790)
    out[:, 0] = np.maximum(out[:, 0], 0)
    pol = out[:, 1] ** 2 + out[:, 2] ** 2 + out[:, 3] ** 2
    with np.errstate(all='ignore'):
        factor = np.where(pol > out[:, 0] ** 2, np.sqrt(out[:, 0] ** 2 / pol), one)
    out[:, 1:] *= factor[:, None]
    return out


def transfer(scratch, cam_map, num, dl_factor, freq_index, nu, dtype):
    """Stokes (I, Q, U, V) x nu^3 at the camera for every ray: scratch (fields, S, R) with S >= max(num), dl_factor (R,)
    = x_unit / (momentum factor x frequency)."""
    xp = dtype
    R = scratch.shape[2]
    s = np.zeros((R, 4), dtype=dtype)
    c0 = F_COEF + 8 * freq_index
    for n in range(int(num.max()) - 1, -1, -1):
        act = n < num
        if not act.any():
            continue
        M = scratch[F_M:F_M + 10, n][:, act].astype(dtype)
        cf = scratch[c0:c0 + 8, n][:, act].astype(dtype)
        dl = scratch[F_DLAM, n][act].astype(dtype) * dl_factor[act].astype(dtype)
        sa = s[act]
        t = np.stack([M[0] * sa[:, 0] + M[1] * sa[:, 1] + M[2] * sa[:, 2], M[3] * sa[:, 0] + M[4] * sa[:, 1] + M[5] * sa[:, 2],
                      M[6] * sa[:, 0] + M[7] * sa[:, 1] + M[8] * sa[:, 2], M[9] * sa[:, 3]], axis=1)
        zero = np.zeros_like(cf[0])
        j = np.stack([cf[0], cf[1], zero, cf[2]], axis=1)
        al = np.stack([cf[3], cf[4], zero, cf[5]], axis=1)
        rho = np.stack([zero, cf[6], zero, cf[7]], axis=1)
        s[act] = couple(t, j, al, rho, dl, xp)
    C = cam_map.astype(dtype)
    nu3 = xp(nu) ** 3
    return np.stack([(C[0] * s[:, 0] + C[1] * s[:, 1] + C[2] * s[:, 2]) * nu3, (C[3] * s[:, 0] + C[4] * s[:, 1] + C[5] * s[:, 2]) * nu3,
                     (C[6] * s[:, 0] + C[7] * s[:, 1] + C[8] * s[:, 2]) * nu3, C[9] * s[:, 3] * nu3], axis=0)
