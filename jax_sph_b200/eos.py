"""Equation-of-state parameter objects, host mirror of jax_sph/eos.py:20-57.

The CUDA kernels evaluate the EoS inline; these classes carry the parameters
(same constructor arguments as the reference) and expose ``p_fn`` / ``rho_fn``
on torch tensors for API compatibility.
"""


class TaitEoS:
    """eos.py:20-38."""

    def __init__(self, p_ref, rho_ref, p_background, gamma):
        self.p_ref, self.rho_ref, self.p_bg, self.gamma = p_ref, rho_ref, p_background, gamma

    def p_fn(self, rho):
        return self.p_ref * ((rho / self.rho_ref) ** self.gamma - 1) + self.p_bg

    def rho_fn(self, p):
        p_temp = p + self.p_ref - self.p_bg
        return self.rho_ref * (p_temp / self.p_ref) ** (1 / self.gamma)


class RIEMANNEoS:
    """eos.py:41-57."""

    def __init__(self, rho_ref, p_background, u_ref):
        self.rho_ref, self.u_ref, self.p_bg = rho_ref, u_ref, p_background

    def p_fn(self, rho):
        return 100 * self.u_ref**2 * (rho - self.rho_ref) + self.p_bg

    def rho_fn(self, p):
        return (p - self.p_bg) / (100 * self.u_ref**2) + self.rho_ref
