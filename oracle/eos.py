"""Equations of state (oracle; test infrastructure only).  jax_sph/eos.py:20-57."""

import numpy as np


class TaitEoS:
    """eos.py:20-38."""

    def __init__(self, p_ref, rho_ref, p_background, gamma):
        self.p_ref = p_ref
        self.rho_ref = rho_ref
        self.p_bg = p_background
        self.gamma = gamma

    def p_fn(self, rho):
        t = rho.dtype.type
        return t(self.p_ref) * ((rho / t(self.rho_ref)) ** t(self.gamma) - t(1)) + t(self.p_bg)

    def rho_fn(self, p):
        t = p.dtype.type
        p_temp = p + t(self.p_ref) - t(self.p_bg)
        return t(self.rho_ref) * (p_temp / t(self.p_ref)) ** t(1 / self.gamma)


class RIEMANNEoS:
    """eos.py:41-57."""

    def __init__(self, rho_ref, p_background, u_ref):
        self.rho_ref = rho_ref
        self.u_ref = u_ref
        self.p_bg = p_background

    def p_fn(self, rho):
        t = rho.dtype.type
        return t(100 * self.u_ref**2) * (rho - t(self.rho_ref)) + t(self.p_bg)

    def rho_fn(self, p):
        t = p.dtype.type
        return (p - t(self.p_bg)) / t(100 * self.u_ref**2) + t(self.rho_ref)
