"""Equation-of-state parameter records for the engine (host mirror of jax_sph/eos.py:20-57).

The pressure is evaluated INSIDE the CUDA sweeps (csrc/phys.cuh, selected by
``SPHB200_EOS_TAIT`` / ``SPHB200_EOS_RIEMANN`` in the engine config); nothing here runs on the
hot path.  The two classes keep the reference's constructor signatures and attribute names
(``solver.WCSPH`` reads ``p_ref``, ``rho_ref``, ``p_bg``, ``gamma`` / ``u_ref`` off them) and
add what the engine needs: ``kind`` (the config enum name) and ``engine_fields()`` (keyword
arguments for ``make_config``).  ``p_fn`` / ``rho_fn`` evaluate the same closed forms on NumPy
arrays or torch tensors, for setting up an initial state (case_setup.py:170 ``eos.p_fn(rho)``).
"""

from typing import Dict


class _EoSRecord:
    """What both laws share: a reference density and a background pressure."""

    kind = ""

    def __init__(self, rho_ref: float, p_background: float):
        if not rho_ref > 0:
            raise ValueError("rho_ref must be positive")
        self.rho_ref = rho_ref
        self.p_bg = p_background

    def engine_fields(self) -> Dict:
        return {"eos": self.kind, "rho_ref": self.rho_ref, "p_bg": self.p_bg}

    def __repr__(self):
        args = ", ".join(f"{k}={v!r}" for k, v in sorted(self.engine_fields().items()))
        return f"{type(self).__name__}({args})"


class TaitEoS(_EoSRecord):
    """Tait law (Adami et al. 2012): stiffness ``p_ref``, exponent ``gamma``;
    reference: jax_sph/eos.py:20-38."""

    kind = "TAIT"

    def __init__(self, p_ref, rho_ref, p_background, gamma):
        super().__init__(rho_ref, p_background)
        self.p_ref, self.gamma = p_ref, gamma

    def engine_fields(self) -> Dict:
        return dict(super().engine_fields(), p_ref=self.p_ref, gamma=self.gamma)

    def p_fn(self, rho):
        compression = (rho / self.rho_ref) ** self.gamma
        return self.p_ref * (compression - 1) + self.p_bg

    def rho_fn(self, p):
        # inverse of p_fn: undo the background shift, then the power law
        gauge = p + self.p_ref - self.p_bg
        return self.rho_ref * (gauge / self.p_ref) ** (1 / self.gamma)


class RIEMANNEoS(_EoSRecord):
    """Linear law of the Riemann solver variant (Zhang, Hu, Adams 2017) with the fixed sound
    speed ``10 u_ref``; reference: jax_sph/eos.py:41-57."""

    kind = "RIEMANN"

    def __init__(self, rho_ref, p_background, u_ref):
        super().__init__(rho_ref, p_background)
        self.u_ref = u_ref

    @property
    def c2(self):
        """Squared artificial sound speed, (10 u_ref)^2, evaluated as the reference does."""
        return 100 * self.u_ref**2

    def engine_fields(self) -> Dict:
        return dict(super().engine_fields(), u_ref=self.u_ref)

    def p_fn(self, rho):
        return self.c2 * (rho - self.rho_ref) + self.p_bg

    def rho_fn(self, p):
        return (p - self.p_bg) / self.c2 + self.rho_ref
