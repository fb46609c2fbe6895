"""WCSPH.forward on an explicit edge list (oracle; test infrastructure only).

Line-by-line NumPy restatement of jax_sph/solver.py for solver in {SPH, RIE, DELTA}
(DELTA is SURVEY.md section 8 row a19, "next": rho_evol_fn_delta solver.py:33-105,
acceleration_delta_fn solver.py:259-313):

* EPS                      solver.py:21
* rho_evol_fn              solver.py:24-30
* rho_evol_riemann_fn      solver.py:108-164
* rho_renorm_fn            solver.py:167-173
* rho_summation_fn         solver.py:176-178
* acceleration_tvf_fn      solver.py:199-213
* tvf_stress_fn            solver.py:216-218
* acceleration_standard_fn solver.py:221-256
* acceleration_fn_riemann  solver.py:316-401
* artificial_viscosity_fn  solver.py:404-428
* gwbc_fn                  solver.py:431-530
* gwbc_fn_riemann_wrapper  solver.py:533-571
* limiter beta_fn          solver.py:574-588
* temperature_derivative   solver.py:591-610
* WCSPH.__init__/forward   solver.py:613-951

Everything is evaluated in the dtype of ``state["r"]`` (float32 reproduces the
reference's x64-disabled run, float64 its default run).  Scatter-adds are
sequential in edge order (``np.add.at``), the order XLA:CPU applies them in;
``fast_segment_sum=True`` switches to ``np.bincount`` (float64 accumulation)
and is only used for the timed CPU baseline.
"""

import numpy as np

from . import kernel as _kernel
from . import space
from .eos import RIEMANNEoS  # noqa: F401  (re-export for callers)

# utils.py:22-32
PAD_VALUE, FLUID, SOLID_WALL, MOVING_WALL, DIRICHLET_WALL = -1, 0, 1, 2, 3
WALL_TAGS = (SOLID_WALL, MOVING_WALL, DIRICHLET_WALL)


def _is_wall(tag):
    return np.isin(tag, WALL_TAGS)


class WCSPH:
    """Mirror of solver.py:613-700 (constructor arguments in the same order)."""

    def __init__(
        self,
        displacement_fn,
        eos,
        g_ext_fn,
        dx,
        dim,
        dt,
        c_ref,
        eta_limiter=3,
        diff_delta=0.02,
        diff_alpha=0.1,
        solver="SPH",
        kernel="QSK",
        h_fac=1.0,
        is_bc_trick=False,
        is_rho_evol=False,
        artificial_alpha=0.0,
        is_free_slip=False,
        is_rho_renorm=False,
        is_heat_conduction=False,
        dtype=np.float64,
        fast_segment_sum=False,
    ):
        if solver not in ("SPH", "RIE", "DELTA"):
            raise NotImplementedError("oracle covers SPH, RIE and DELTA")
        self.diff_delta = diff_delta
        self.diff_alpha = diff_alpha
        self.h = h_fac * dx  # `support` of the DELTA wrappers (solver.py:678, :683)
        self.displacement_fn = displacement_fn
        self.eos = eos
        self.g_ext_fn = g_ext_fn
        self.dx = dx
        self.dim = dim
        self.dt = dt
        self.c_ref = c_ref
        self.eta_limiter = eta_limiter
        self.solver = solver
        self.is_bc_trick = is_bc_trick
        self.is_rho_evol = is_rho_evol
        self.artificial_alpha = artificial_alpha
        self.is_free_slip = is_free_slip
        self.is_rho_renorm = is_rho_renorm
        self.is_heat_conduction = is_heat_conduction
        self.dtype = np.dtype(dtype)
        self.fast = fast_segment_sum
        self._kernel_fn = _kernel.KERNELS[kernel](h=h_fac * dx, dim=dim, dtype=dtype)
        self.EPS = self.dtype.type(np.finfo(self.dtype).eps)

    # -- helpers ---------------------------------------------------------
    def _seg(self, data, seg, N):
        """ops.segment_sum(data, seg, N)."""
        if self.fast:
            if data.ndim == 1:
                return np.bincount(seg, weights=data, minlength=N).astype(data.dtype)
            return np.stack(
                [np.bincount(seg, weights=data[:, k], minlength=N) for k in range(data.shape[1])],
                axis=1,
            ).astype(data.dtype)
        out = np.zeros((N,) + data.shape[1:], dtype=data.dtype)
        np.add.at(out, seg, data)
        return out

    @staticmethod
    def _dot(a, b):
        acc = a[:, 0] * b[:, 0]
        for k in range(1, a.shape[1]):
            acc = acc + a[:, k] * b[:, k]
        return acc

    # -- forward ---------------------------------------------------------
    def forward_wrapper(self):
        return self.forward

    def forward(self, state, idx):
        """solver.py:705-949.  ``idx`` is the (2, E) edge list (padding allowed)."""
        t = self.dtype.type
        EPS = self.EPS
        kf = self._kernel_fn
        seg = self._seg
        dot = self._dot

        r, tag, mass, eta = state["r"], state["tag"], state["mass"], state["eta"]
        u, v, dudt, dvdt = state["u"], state["v"], state["dudt"], state["dvdt"]
        rho, drhodt, p = state["rho"], state["drhodt"], state["p"]
        nw, kappa, Cp = state["nw"], state["kappa"], state["Cp"]
        temperature, dTdt = state["T"], state["dTdt"]
        N = len(r)

        idx = np.asarray(idx)
        real = idx[0] < N  # padding edges (N, N) contribute nothing (clamp + drop)
        i_s = idx[0][real].astype(np.int64)
        j_s = idx[1][real].astype(np.int64)

        dr_i_j = self.displacement_fn(r[i_s], r[j_s])  # :723-724
        dist = space.distance(dr_i_j)  # :725
        w_dist = kf.w(dist)  # :726
        e_s = dr_i_j / (dist[:, None] + EPS)  # :728
        grad_w_dist_norm = kf.grad_w(dist)  # :730
        grad_w_dist = grad_w_dist_norm[:, None] * e_s  # :731

        g_ext = self.g_ext_fn(r).astype(self.dtype)  # :734
        wall_mask = np.where(_is_wall(tag), t(1.0), t(0.0))  # :737
        fluid_mask = np.where(tag == FLUID, t(1.0), t(0.0))  # :738

        # Riemann velocity BCs :741-744 (riemann_velocities :547-552)
        if self.is_bc_trick and self.solver == "RIE" and not self.is_free_slip:
            w_dist_fluid = w_dist * fluid_mask[j_s]
            u_wall_nom = seg(w_dist_fluid[:, None] * u[j_s], i_s, N)
            u_wall_denom = seg(w_dist_fluid, i_s, N)
            u_tilde = u_wall_nom / (u_wall_denom[:, None] + EPS)
        else:
            u_tilde = u

        # density :750-796
        dt_s = t(self.dt)
        if self.is_rho_evol and self.solver == "SPH":
            v_j_s = (mass / rho)[j_s]  # :26
            temp = v_j_s * ((u[i_s] - u[j_s]) * grad_w_dist).sum(axis=1)  # :27
            drhodt = rho * seg(temp, i_s, N)  # :28
            rho = rho + dt_s * drhodt  # :29
            if self.is_rho_renorm:
                rho = self._rho_renorm(rho, mass, i_s, j_s, w_dist, N)
        elif self.is_rho_evol and self.solver == "DELTA":  # :757-771
            rho, drhodt = self._rho_evol_delta(
                rho, mass, dr_i_j, dist, u, i_s, j_s, dt_s, N, fluid_mask[j_s]
            )
            if self.is_rho_renorm:
                rho = self._rho_renorm(rho, mass, i_s, j_s, w_dist, N)
        elif self.is_rho_evol and self.solver == "RIE":
            temp = self._rho_evol_riemann(
                e_s, rho[i_s], rho[j_s], mass[j_s], u[i_s], u[j_s], p[i_s], p[j_s],
                dr_i_j, dist, wall_mask[j_s], nw[j_s], g_ext[i_s],
            )
            drhodt = seg(temp, i_s, N) * fluid_mask  # :789
            rho = rho + dt_s * drhodt  # :790
            if self.is_rho_renorm:
                rho = self._rho_renorm(rho, mass, i_s, j_s, w_dist, N)
        else:
            rho_ = mass * seg(w_dist, i_s, N)  # :176-178, :795
            rho = np.where(fluid_mask.astype(bool), rho_, rho)  # :796

        p = self.eos.p_fn(rho)  # :801
        background_pressure_tvf = self.eos.p_fn(np.zeros_like(p))  # :802

        # wall boundary conditions :806-828
        if self.is_bc_trick and self.solver in ("SPH", "DELTA"):  # :806
            p, rho, u, v, temperature = self._gwbc(
                temperature, rho, tag, u, v, p, g_ext, i_s, j_s, w_dist, dr_i_j, nw, N
            )
            mask = None
        elif self.is_bc_trick and self.solver == "RIE":
            if self.is_free_slip:
                mask = fluid_mask[i_s]  # free_weight :537-538
            else:
                mask = np.ones_like(tag[i_s]).astype(self.dtype)  # :544-545
            if self.is_heat_conduction:  # heat_bc :556-565
                w_j_s_fluid = w_dist * fluid_mask[j_s]
                w_i_sum_wf = seg(w_j_s_fluid, i_s, N)
                t_wall_unnorm = seg(w_j_s_fluid * temperature[j_s], i_s, N)
                t_wall = t_wall_unnorm / (w_i_sum_wf + EPS)
                m_ = np.isin(tag, (SOLID_WALL, MOVING_WALL))
                temperature = np.where(m_, t_wall, temperature)
        elif self.solver == "RIE":
            mask = np.ones_like(tag[i_s]).astype(self.dtype)  # :828
        else:
            mask = None

        # heat conduction :832-850
        if self.is_heat_conduction:
            temperature = temperature + dt_s * dTdt  # :834
            out = self._temperature_derivative(
                e_s, dr_i_j, dist, rho[i_s], rho[j_s], mass[j_s], kappa[i_s], kappa[j_s],
                Cp[i_s], temperature[i_s], temperature[j_s],
            )
            dTdt = seg(out, i_s, N)  # :850

        # momentum :854-910
        if self.solver == "SPH":
            out = self._acceleration_standard(
                dr_i_j, dist, rho[i_s], rho[j_s], u[i_s], u[j_s], v[i_s], v[j_s],
                mass[i_s], mass[j_s], eta[i_s], eta[j_s], p[i_s], p[j_s],
            )
        elif self.solver == "DELTA":  # :871-888
            out = self._acceleration_standard(
                dr_i_j, dist, rho[i_s], rho[j_s], u[i_s], u[j_s], v[i_s], v[j_s],
                mass[i_s], mass[j_s], eta[i_s], eta[j_s], p[i_s], p[j_s],
            ) + self._acceleration_delta_diff(
                dr_i_j, dist, rho[i_s], rho[j_s], u[i_s], u[j_s], mass[j_s], fluid_mask[j_s]
            )
        else:
            out = self._acceleration_riemann(
                e_s, dr_i_j, dist, rho[i_s], rho[j_s], mass[j_s], mass[i_s], u[i_s], u[j_s],
                p[i_s], p[j_s], eta[i_s], eta[j_s], wall_mask[j_s], mask, nw[j_s],
                g_ext[i_s], u_tilde[j_s],
            )
        dudt = seg(out, i_s, N)  # :910

        out_tv = self._acceleration_tvf(
            dr_i_j, dist, rho[i_s], rho[j_s], mass[i_s], mass[j_s], background_pressure_tvf[i_s]
        )  # :912-920
        dvdt = seg(out_tv, i_s, N)  # :921

        if self.artificial_alpha != 0.0:  # :925-928
            dudt = dudt + self._artificial_viscosity(
                rho, mass, u, tag, i_s, j_s, dr_i_j, dist, grad_w_dist, N
            )

        return {
            "r": r,
            "tag": tag,
            "u": u,
            "v": v,
            "drhodt": drhodt,
            "dudt": dudt + g_ext,  # :936
            "dvdt": dvdt,
            "rho": rho,
            "p": p,
            "mass": mass,
            "eta": eta,
            "dTdt": dTdt,
            "T": temperature,
            "kappa": kappa,
            "Cp": Cp,
            "nw": nw,
        }

    # -- pieces ----------------------------------------------------------
    def _rho_renorm(self, rho, mass, i_s, j_s, w_dist, N):
        """solver.py:167-173."""
        t = self.dtype.type
        nominator = self._seg(mass[j_s] * w_dist, i_s, N)
        den = self._seg((mass / rho)[j_s] * w_dist, i_s, N)
        den = np.where(den > 1, t(1), den)
        return nominator / den

    def _riemann_states(self, e_ij, rho_i, rho_j, u_i, u_j, p_i, p_j, r_ij, wall_j, n_w_j, g_i):
        """Shared part of solver.py:134-150 and :346-362."""
        t = self.dtype.type
        dot = self._dot
        is_w = wall_j == t(1.0)  # jnp.isin(wall_mask_j, wall_tags) on a 0/1 float mask
        u_L = np.where(is_w, dot(u_i, -n_w_j), dot(u_i, -e_ij))
        p_L = p_i
        rho_L = rho_i
        u_R = np.where(is_w, -u_L + t(2) * dot(u_j, n_w_j), dot(u_j, -e_ij))
        p_R = np.where(is_w, p_L + rho_L * dot(g_i, -r_ij), p_j)
        rho_R = np.where(is_w, self.eos.rho_fn(p_R), rho_j)
        return is_w, u_L, p_L, rho_L, u_R, p_R, rho_R

    def _rho_evol_riemann(
        self, e_s, rho_i, rho_j, m_j, u_i, u_j, p_i, p_j, r_ij, d_ij, wall_j, n_w_j, g_i
    ):
        """solver.py:111-162."""
        t = self.dtype.type
        e_ij = e_s
        kernel_grad = self._kernel_fn.grad_w(d_ij)[:, None] * e_ij
        _, u_L, p_L, rho_L, u_R, p_R, rho_R = self._riemann_states(
            e_ij, rho_i, rho_j, u_i, u_j, p_i, p_j, r_ij, wall_j, n_w_j, g_i
        )
        U_avg = (u_L + u_R) / t(2)
        v_avg = (u_i + u_j) / t(2)
        rho_avg = (rho_L + rho_R) / t(2)
        U_star = U_avg + t(0.5) * (p_L - p_R) / (rho_avg * t(self.c_ref))
        v_star = U_star[:, None] * (-e_ij) + (v_avg - U_avg[:, None] * (-e_ij))
        return t(2) * rho_i * m_j / rho_j * self._dot(u_i - v_star, kernel_grad)

    def _beta(self, u_L, u_R):
        """solver.py:574-588."""
        t = self.dtype.type
        if self.eta_limiter == -1:
            return t(self.c_ref)
        temp = t(self.eta_limiter) * np.maximum(u_L - u_R, np.zeros_like(u_L))
        return np.minimum(temp, np.full_like(temp, t(self.c_ref)))

    def _acceleration_riemann(
        self, e_s, r_ij, d_ij, rho_i, rho_j, m_j, m_i, u_i, u_j, p_i, p_j, eta_i, eta_j,
        wall_j, mask, n_w_j, g_i, u_tilde_j,
    ):
        """solver.py:319-399."""
        t = self.dtype.type
        EPS = self.EPS
        e_ij = e_s
        kernel_part_diff = self._kernel_fn.grad_w(d_ij)
        kernel_grad = kernel_part_diff[:, None] * e_ij
        is_w, u_L, p_L, rho_L, u_R, p_R, rho_R = self._riemann_states(
            e_ij, rho_i, rho_j, u_i, u_j, p_i, p_j, r_ij, wall_j, n_w_j, g_i
        )
        P_avg = (p_L + p_R) / t(2)
        rho_avg = (rho_L + rho_R) / t(2)
        eta_ij = t(2) * eta_i * eta_j / (eta_i + eta_j + EPS)
        P_star = P_avg + t(0.5) * rho_avg * (u_L - u_R) * self._beta(u_L, u_R)
        eq_9 = (t(-2) * m_j * (P_star / (rho_i * rho_j)))[:, None] * kernel_grad
        u_d = t(2) * u_j - u_tilde_j
        v_ij = np.where(is_w[:, None], u_i - u_d, u_i - u_j)
        eq_6 = (t(2) * m_j * eta_ij / (rho_i * rho_j))[:, None] * v_ij / (d_ij + EPS)[:, None]
        eq_6 = eq_6 * (kernel_part_diff * mask)[:, None]
        wv = ((m_i / rho_i) ** 2 + (m_j / rho_j) ** 2) / m_i
        c = wv * kernel_part_diff / (d_ij + EPS)
        # tvf_stress_fn(rho, u, u) = outer(rho*u, u - u) == 0, kept for fidelity
        zero = u_i - u_i
        A_r_fluid = ((rho_i[:, None] * u_i) * self._dot(zero, r_ij)[:, None]
                     + (rho_j[:, None] * u_j) * self._dot(u_j - u_j, r_ij)[:, None]) / t(2)
        A_r_wall = ((rho_i[:, None] * u_i) * self._dot(zero, r_ij)[:, None]
                    + (rho_j[:, None] * u_d) * self._dot(u_d - u_d, r_ij)[:, None]) / t(2)
        a_eq_8 = c[:, None] * np.where(is_w[:, None], A_r_wall, A_r_fluid)
        return eq_9 + eq_6 + a_eq_8

    def _acceleration_standard(
        self, r_ij, d_ij, rho_i, rho_j, u_i, u_j, v_i, v_j, m_i, m_j, eta_i, eta_j, p_i, p_j
    ):
        """solver.py:224-254."""
        t = self.dtype.type
        EPS = self.EPS
        eta_ij = t(2) * eta_i * eta_j / (eta_i + eta_j + EPS)
        p_ij = (rho_j * p_i + rho_i * p_j) / (rho_i + rho_j)
        wv = ((m_i / rho_i) ** 2 + (m_j / rho_j) ** 2) / m_i
        c = wv * self._kernel_fn.grad_w(d_ij) / (d_ij + EPS)
        # dot(outer(rho*u, v-u), r) = rho*u * ((v-u).r)
        A_r = (
            (rho_i[:, None] * u_i) * self._dot(v_i - u_i, r_ij)[:, None]
            + (rho_j[:, None] * u_j) * self._dot(v_j - u_j, r_ij)[:, None]
        ) / t(2)
        u_ij = u_i - u_j
        return c[:, None] * (-p_ij[:, None] * r_ij + A_r + eta_ij[:, None] * u_ij)

    def _acceleration_delta_diff(self, r_ij, d_ij, rho_i, rho_j, u_i, u_j, m_j, fluidmask_j):
        """The Delta-SPH velocity diffusion of acceleration_delta_fn, solver.py:297-311 (a_eq_8,
        :280-295, is the standard acceleration)."""
        t = self.dtype.type
        EPS = self.EPS
        e_ij = r_ij / (d_ij + EPS)[:, None]
        kernel_grad = self._kernel_fn.grad_w(d_ij)[:, None] * e_ij
        V_j = m_j / rho_j
        pi_ij = self._dot(u_j - u_i, -r_ij) / (d_ij + EPS) ** 2
        rho_ref = t(self.eos.rho_ref)
        # V_j * pi_ij * kernel_grad * alpha * support * c_ref * rho_ref / rho_i * fluidmask_j
        out = (V_j * pi_ij)[:, None] * kernel_grad
        out = out * t(self.diff_alpha) * t(self.h) * t(self.c_ref) * rho_ref
        return out / rho_i[:, None] * fluidmask_j[:, None]

    def _rho_evol_delta(self, rho, mass, r_ij, d_ij, u, i_s, j_s, dt, N, fluidmask_j):
        """rho_evol_fn_delta, solver.py:36-103 (Marrone et al. 2011)."""
        t = self.dtype.type
        EPS = self.EPS
        seg = self._seg
        d = r_ij.shape[1]
        e_ab = r_ij / (d_ij + EPS)[:, None]  # :42
        kernel_grad = self._kernel_fn.grad_w(d_ij)[:, None] * e_ab  # :43
        V_i, V_j = (mass / rho)[i_s], (mass / rho)[j_s]  # :44-45
        # :52-57  L = inv(sum tensordot(-r_ab, kernel_grad * V_b))
        temp = (-r_ij)[:, :, None] * (kernel_grad * V_j[:, None])[:, None, :]
        L_mati = np.linalg.inv(seg(temp.reshape(len(i_s), d * d), i_s, N).reshape(N, d, d))
        temp = (r_ij)[:, :, None] * (kernel_grad * V_i[:, None])[:, None, :]
        L_matj = np.linalg.inv(seg(temp.reshape(len(i_s), d * d), j_s, N).reshape(N, d, d))
        # :59-65
        gi = np.einsum("eab,eb->ea", L_mati[i_s], kernel_grad * V_j[:, None])
        rho_grad_term_i = seg((rho[j_s] - rho[i_s])[:, None] * gi, i_s, N)
        gj = np.einsum("eab,eb->ea", L_matj[j_s], -kernel_grad * V_i[:, None])
        rho_grad_term_j = seg((rho[i_s] - rho[j_s])[:, None] * gj, i_s, N)
        # :67-93
        rho_term = (t(2) * (rho[j_s] - rho[i_s]))[:, None] * (-r_ij) / ((d_ij + EPS) ** 2)[:, None]
        psi_ij = rho_term - rho_grad_term_i[i_s] - rho_grad_term_j[j_s]
        diff_term = seg(self._dot(psi_ij, kernel_grad) * V_j * fluidmask_j, i_s, N)
        # :95-103
        cont = seg(self._dot(u[i_s] - u[j_s], kernel_grad) * V_j, i_s, N)
        drhodt = rho * cont + t(self.c_ref) * t(self.diff_delta) * t(self.h) * diff_term
        return (rho + dt * drhodt).astype(self.dtype), drhodt.astype(self.dtype)

    def _acceleration_tvf(self, r_ij, d_ij, rho_i, rho_j, m_i, m_j, p_bg_i):
        """solver.py:202-211."""
        wv = ((m_i / rho_i) ** 2 + (m_j / rho_j) ** 2) / m_i
        c = wv * self._kernel_fn.grad_w(d_ij) / (d_ij + self.EPS)
        return (c * self.dtype.type(1.0) * p_bg_i)[:, None] * r_ij

    def _artificial_viscosity(self, rho, mass, u, tag, i_s, j_s, dr_i_j, dist, grad_w_dist, N):
        """solver.py:404-428 (u_ref hard-coded to 1.0 -> c_ab = 10)."""
        t = self.dtype.type
        h_ab = self.dx
        c_ab = 10.0 * 1.0
        rho_ab = (rho[i_s] + rho[j_s]) / t(2)
        numerator = mass[j_s] * t(self.artificial_alpha * h_ab * c_ab)
        numerator = numerator * ((u[i_s] - u[j_s]) * dr_i_j).sum(axis=1)
        numerator = numerator[:, None] * grad_w_dist
        denominator = (rho_ab * (dist**2 + t(0.01 * h_ab**2)))[:, None]
        mask_fluid = tag == FLUID
        mfe = (mask_fluid[j_s] * mask_fluid[i_s]).astype(self.dtype)
        res = mfe[:, None] * numerator / denominator
        return self._seg(res, i_s, N)

    def _gwbc(self, temperature, rho, tag, u, v, p, g_ext, i_s, j_s, w_dist, dr_i_j, nw, N):
        """solver.py:450-528."""
        t = self.dtype.type
        EPS = self.EPS
        seg = self._seg
        mask_bc = _is_wall(tag)
        mask_j_s_fluid = np.where(tag[j_s] == FLUID, t(1.0), t(0.0))
        w_j_s_fluid = w_dist * mask_j_s_fluid
        w_i_sum_wf = seg(w_j_s_fluid, i_s, N)

        def no_slip(x):
            x_wall_unnorm = seg(w_j_s_fluid[:, None] * x[j_s], i_s, N)
            x_wall = x_wall_unnorm / (w_i_sum_wf[:, None] + EPS)
            return np.where(mask_bc[:, None], t(2) * x - x_wall, x)

        def free_slip(x, win):
            x_wall_unnorm = seg(w_j_s_fluid[:, None] * x[j_s], i_s, N)
            x_wall = x_wall_unnorm / (w_i_sum_wf[:, None] + EPS)
            x_wall = win * (x_wall * win).sum(axis=1, keepdims=True)
            return np.where(mask_bc[:, None], t(2) * x - x_wall, x)

        if self.is_free_slip:
            u = free_slip(u, -nw)
            v = free_slip(v, -nw)
        else:
            u = no_slip(u)
            v = no_slip(v)

        p_wall_unnorm = seg(w_j_s_fluid * p[j_s], i_s, N)
        rho_wf_sum = (rho[j_s] * w_j_s_fluid)[:, None] * dr_i_j
        rho_wf_sum = seg(rho_wf_sum, i_s, N)
        p_wall_ext = (g_ext * rho_wf_sum).sum(axis=1)
        p_wall = (p_wall_unnorm + p_wall_ext) / (w_i_sum_wf + EPS)
        p = np.where(mask_bc, p_wall, p)
        rho = self.eos.rho_fn(p)  # :516, all particles

        if self.is_heat_conduction:
            t_wall_unnorm = seg(w_j_s_fluid * temperature[j_s], i_s, N)
            t_wall = t_wall_unnorm / (w_i_sum_wf + EPS)
            m_ = np.isin(tag, (SOLID_WALL, MOVING_WALL))
            temperature = np.where(m_, t_wall, temperature)
        return p, rho, u, v, temperature

    def _temperature_derivative(
        self, e_s, r_ij, d_ij, rho_i, rho_j, m_j, kappa_i, kappa_j, Cp_i, T_i, T_j
    ):
        """solver.py:594-608."""
        t = self.dtype.type
        kg = self._kernel_fn.grad_w(d_ij)[:, None] * e_s
        with np.errstate(invalid="ignore", divide="ignore"):
            eff = (kappa_i * kappa_j) / (kappa_i + kappa_j)
            F_ab = self._dot(r_ij, kg) / ((d_ij * d_ij) + self.EPS)
            return (t(4) * m_j * eff * (T_i - T_j) * F_ab) / (Cp_i * rho_i * rho_j)
