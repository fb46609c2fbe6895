"""Equations of state of the oracle (test infrastructure only; never on the product path).

Restates jax_sph/eos.py:20-57 for NumPy arrays of a given dtype: every Python scalar of the
reference's expressions is cast to the array's dtype first, which is what jax's weak typing does
to them, so that the float32 oracle rounds where the float32 reference rounds.  The closed forms
live in four module-level functions; the two classes only hold the parameters under the
reference's attribute names (the solver restatement reads them) and forward to the functions.
"""


def _cast(array):
    """dtype constructor of `array`: scalars pass through it before they meet the array."""
    return array.dtype.type


def tait_p(rho, p_ref, rho_ref, p_bg, gamma):
    """eos.py:32-33  p = p_ref ((rho / rho_ref)^gamma - 1) + p_bg."""
    c = _cast(rho)
    return c(p_ref) * ((rho / c(rho_ref)) ** c(gamma) - c(1)) + c(p_bg)


def tait_rho(p, p_ref, rho_ref, p_bg, gamma):
    """eos.py:35-37  the inverse of tait_p."""
    c = _cast(p)
    shifted = p + c(p_ref) - c(p_bg)
    return c(rho_ref) * (shifted / c(p_ref)) ** c(1 / gamma)


def riemann_p(rho, rho_ref, p_bg, u_ref):
    """eos.py:53-54  p = 100 u_ref^2 (rho - rho_ref) + p_bg."""
    c = _cast(rho)
    return c(100 * u_ref**2) * (rho - c(rho_ref)) + c(p_bg)


def riemann_rho(p, rho_ref, p_bg, u_ref):
    """eos.py:56-57  the inverse of riemann_p."""
    c = _cast(p)
    return (p - c(p_bg)) / c(100 * u_ref**2) + c(rho_ref)


class TaitEoS:
    """Parameters of eos.py:20-38 (Adami et al. 2012)."""

    def __init__(self, p_ref, rho_ref, p_background, gamma):
        self.p_ref, self.rho_ref, self.p_bg, self.gamma = p_ref, rho_ref, p_background, gamma

    def p_fn(self, rho):
        return tait_p(rho, self.p_ref, self.rho_ref, self.p_bg, self.gamma)

    def rho_fn(self, p):
        return tait_rho(p, self.p_ref, self.rho_ref, self.p_bg, self.gamma)


class RIEMANNEoS:
    """Parameters of eos.py:41-57 (Zhang, Hu, Adams 2017)."""

    def __init__(self, rho_ref, p_background, u_ref):
        self.rho_ref, self.p_bg, self.u_ref = rho_ref, p_background, u_ref

    def p_fn(self, rho):
        return riemann_p(rho, self.rho_ref, self.p_bg, self.u_ref)

    def rho_fn(self, p):
        return riemann_rho(p, self.rho_ref, self.p_bg, self.u_ref)
