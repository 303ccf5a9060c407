"""Minimal particle table (EvtGen names -> mass, width in MeV) used by ``GenMultiDecay.from_dict``
when the ``particle`` package (the reference's source, genmultidecay.py:244,298,319) is absent.
Values: PDG averages; widths of long-lived particles from hbar / lifetime."""

_HBAR = 6.582119569e-22  # MeV s

_TABLE = {
    # name: (mass, width)
    "gamma": (0.0, 0.0),
    "e-": (0.51099895, 0.0),
    "mu-": (105.6583755, _HBAR / 2.1969811e-6),
    "tau-": (1776.86, _HBAR / 2.903e-13),
    "nu_e": (0.0, 0.0),
    "nu_mu": (0.0, 0.0),
    "nu_tau": (0.0, 0.0),
    "pi+": (139.57039, _HBAR / 2.6033e-8),
    "pi0": (134.9768, _HBAR / 8.43e-17),
    "K+": (493.677, _HBAR / 1.2380e-8),
    "K0": (497.611, 0.0),
    "K_S0": (497.611, _HBAR / 8.954e-11),
    "K_L0": (497.611, _HBAR / 5.116e-8),
    "eta": (547.862, 1.31e-3),
    "rho0": (775.26, 147.4),
    "rho+": (775.11, 149.1),
    "omega": (782.66, 8.68),
    "K*0": (895.55, 47.3),
    "K*+": (891.67, 51.4),
    "eta'": (957.78, 0.188),
    "phi": (1019.461, 4.249),
    "p+": (938.27208816, 0.0),
    "n0": (939.5654205, _HBAR / 878.4),
    "K_1+": (1253.0, 90.0),
    "K_10": (1253.0, 90.0),
    "D0": (1864.84, _HBAR / 4.101e-13),
    "D+": (1869.66, _HBAR / 1.033e-12),
    "D_s+": (1968.35, _HBAR / 5.04e-13),
    "D*0": (2006.85, 0.0555),
    "D*+": (2010.26, 0.0834),
    "J/psi": (3096.9, 0.0926),
    "psi(2S)": (3686.1, 0.294),
    "B0": (5279.66, _HBAR / 1.519e-12),
    "B+": (5279.34, _HBAR / 1.638e-12),
    "B_s0": (5366.92, _HBAR / 1.520e-12),
    "Lambda_b0": (5619.6, _HBAR / 1.471e-12),
    "Upsilon": (9460.3, 0.05402),
}

_CONJUGATE = {
    "e-": "e+", "mu-": "mu+", "tau-": "tau+", "pi+": "pi-", "K+": "K-", "K0": "anti-K0", "rho+": "rho-",
    "K*0": "anti-K*0", "K*+": "K*-", "p+": "anti-p-", "n0": "anti-n0", "K_1+": "K_1-", "K_10": "anti-K_10",
    "D0": "anti-D0", "D+": "D-", "D_s+": "D_s-", "D*0": "anti-D*0", "D*+": "D*-", "B0": "anti-B0", "B+": "B-",
    "B_s0": "anti-B_s0", "Lambda_b0": "anti-Lambda_b0", "nu_e": "anti-nu_e", "nu_mu": "anti-nu_mu",
    "nu_tau": "anti-nu_tau",
}
for _name, _anti in _CONJUGATE.items():
    _TABLE[_anti] = _TABLE[_name]


class _Entry:
    def __init__(self, name, mass, width):
        self.name, self.mass, self.width = name, mass, width


def lookup(name: str):
    """``Particle.from_evtgen_name(name)`` stand-in: object with ``.mass`` and ``.width`` (MeV)."""
    try:
        from particle import Particle  # the reference's source of masses and widths

        part = Particle.from_evtgen_name(name)
        return _Entry(name, part.mass, part.width if part.width is not None else 0.0)
    except ImportError:
        pass
    if name not in _TABLE:
        raise KeyError(f"particle '{name}' is not in the built-in table; install the `particle` package")
    return _Entry(name, *_TABLE[name])
