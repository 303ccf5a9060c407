#!/usr/bin/env python3
"""Writes tests/golden/reference_vectors.npz: outputs of the REFERENCE'S OWN SOURCE CODE on fixed uniform draws.

The reference (/root/reference/src/phasespace) imports TensorFlow, which is not installable here.  But the only
TensorFlow surface its generator touches is ``tensorflow.experimental.numpy`` -- by construction a NumPy-compatible
API -- plus ``tf.function`` (a decorator), ``tf.random.Generator.uniform``, ``tf.Variable`` (isinstance checks),
``tf.shape`` and ``tf.debugging.assert_greater_equal``.  This script installs a stand-in ``tensorflow`` module that maps
``tensorflow.experimental.numpy`` to NumPy itself and the rest to a few lines each, then imports and runs the reference's
unmodified ``GenParticle.generate`` (phasespace.py:628-717 and everything below it).  The random generator handed to it
serves its ``rng.uniform((n, k))`` calls from consecutive columns of a fixed table, i.e. the table IS the reference's
draw order.  What is recorded is therefore the reference's algorithm -- every slice, concatenate, rotation and boost in
its own operation order -- evaluated in IEEE fp64 by NumPy instead of by TensorFlow's CPU kernels (which may differ in
the last bit of sqrt / sin / cos and in the association of a few reductions).

Run here (needs /root/reference; the GPU box does not have it, the tests only read the .npz):

    python tests/golden/make_reference_golden.py
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REFERENCE_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(__file__), "reference_vectors.npz")

B0_MASS, PION_MASS, KAON_MASS = 5279.58, 139.57018, 493.677  # tests/helpers/decays.py:14-20 of the reference
KSTARZ_MASS, K1_MASS = 895.81, 1272.0


class TableGenerator:
    """Stands in for tf.random.Generator: uniform((n, k)) returns the next k columns of a fixed (n, n_draws) table."""

    def __init__(self, table):
        self.table = np.asarray(table, dtype=np.float64)
        self.col = 0

    @classmethod
    def from_seed(cls, seed):
        raise RuntimeError("the golden script always passes a generator")

    def uniform(self, shape, dtype=None, **_):
        n, k = int(shape[0]), int(shape[1])
        assert n == self.table.shape[0], (n, self.table.shape)
        out = self.table[:, self.col:self.col + k].copy()
        assert out.shape == (n, k), "uniform table exhausted"
        self.col += k
        return out


def install_tensorflow_stand_in():
    tf = types.ModuleType("tensorflow")
    experimental = types.ModuleType("tensorflow.experimental")
    tf.experimental = experimental
    experimental.numpy = np  # tensorflow.experimental.numpy IS the NumPy API

    def function(*args, **kwargs):  # tf.function(...)(f) and @tf.function
        if args and callable(args[0]) and not kwargs:
            return args[0]
        return lambda f: f

    def assert_greater_equal(x, y, message="", name=None):
        if not np.all(np.asarray(x) >= np.asarray(y)):
            raise tf.errors.InvalidArgumentError(message)

    tf.function = function
    tf.Variable = type("Variable", (), {})
    tf.Tensor = np.ndarray
    tf.float64 = np.float64
    tf.shape = lambda x: np.asarray(np.shape(x))
    # get_mass() asks for the global generator even when the mass function takes no seed (phasespace.py:176)
    global_generator = TableGenerator(np.empty((0, 0)))
    tf.random = types.SimpleNamespace(Generator=TableGenerator, get_global_generator=lambda: global_generator)
    tf.errors = types.SimpleNamespace(InvalidArgumentError=type("InvalidArgumentError", (Exception,), {}))
    tf.debugging = types.SimpleNamespace(assert_greater_equal=assert_greater_equal)
    tf.config = types.SimpleNamespace(run_functions_eagerly=lambda flag: None)
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.experimental"] = experimental
    sys.modules["tensorflow.experimental.numpy"] = np
    try:
        import vector  # noqa: F401  (phasespace.py:676-680 imports it whenever boost_to is given)
    except ImportError:
        vec = types.ModuleType("vector")
        vec.Vector = type("Vector", (), {})
        vec.Momentum = type("Momentum", (), {})
        vec.Lorentz = type("Lorentz", (), {})
        sys.modules["vector"] = vec


def recording_mass(store, name, frac):
    """A resonance mass function ``f(min_mass, max_mass, n_events)`` (phasespace.py:144-184): a fixed per-event
    fraction of the allowed range; what it returned is recorded so that the oracle / kernel can be fed the same masses."""

    def mass(min_mass, max_mass, n_events):
        out = min_mass + frac * (max_mass - min_mass)
        store[name] = np.array(out, dtype=np.float64).reshape(-1)
        return out

    return mass


def boost_rows(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.normal(0.0, 20000.0, (n, 3))
    return np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + B0_MASS**2)], axis=1)


def n_draws_flat(n_body):
    return 3 * n_body - 4


def build(ref):
    out = {}
    rng = np.random.default_rng(20261020)

    def run(name, decay, n, n_draws, boost=None, masses_store=None):
        table = rng.random((n, n_draws))
        gen = TableGenerator(table)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # the rest-frame boost evaluates 0/0 before tnp.where discards it
            weights, weights_max, parts = decay.generate(n, boost_to=boost, normalize_weights=False, seed=gen)
        assert gen.col == n_draws, (name, gen.col, n_draws)
        out[f"{name}/draws"] = table
        if boost is not None:
            out[f"{name}/boost"] = np.asarray(boost)
        out[f"{name}/weights"] = np.asarray(weights)
        out[f"{name}/weights_max"] = np.asarray(weights_max)
        out[f"{name}/names"] = np.array(list(parts.keys()))
        out[f"{name}/particles"] = np.stack([np.asarray(v) for v in parts.values()])
        for key, val in (masses_store or {}).items():
            out[f"{name}/user_mass/{key}"] = val

    n = 256
    for nb in (2, 3, 4, 10):
        run(f"flat{nb}", ref.nbody_decay(B0_MASS, [PION_MASS] * nb), n, n_draws_flat(nb))
    run("flat5_mixed", ref.nbody_decay(B0_MASS, [KAON_MASS, PION_MASS, 0.0, 105.6583755, PION_MASS]), n, n_draws_flat(5))
    run("flat4_boost", ref.nbody_decay(B0_MASS, [PION_MASS] * 4), n, n_draws_flat(4), boost=boost_rows(n, 7))
    run("flat3_boost_one_row", ref.nbody_decay(B0_MASS, [PION_MASS] * 3), n, n_draws_flat(3), boost=boost_rows(1, 8))

    # B0 -> K*0(-> K+ pi-) gamma with a fixed-mass K*0 (README.rst:115-183) and with a callable K*0 mass
    def kstar_gamma(kstar_mass):
        return ref.GenParticle("B0", B0_MASS).set_children(
            ref.GenParticle("K*0", kstar_mass).set_children(ref.GenParticle("K+", KAON_MASS), ref.GenParticle("pi-", PION_MASS)),
            ref.GenParticle("gamma", 0.0))

    run("kstar_gamma_fixed", kstar_gamma(KSTARZ_MASS), n, 4)
    store = {}
    frac = np.clip(rng.beta(2.0, 30.0, n), 1e-6, 1.0)  # a resonance-like shape hugging the lower limit
    run("kstar_gamma_callable", kstar_gamma(recording_mass(store, "K*0", frac)), n, 4, masses_store=store)
    store = {}
    run("kstar_gamma_callable_boost", kstar_gamma(recording_mass(store, "K*0", frac)), n, 4, boost=boost_rows(n, 9), masses_store=store)

    # B+ -> K1(-> K*0(-> K+ pi-) pi+) gamma, tests/helpers/decays.py:47-76, both resonances callable
    store = {}
    f1 = np.clip(rng.beta(2.0, 12.0, n), 1e-6, 1.0)
    f2 = np.clip(rng.beta(2.0, 6.0, n), 1e-6, 1.0)
    k1 = ref.GenParticle("B+", B0_MASS).set_children(
        ref.GenParticle("K1+", recording_mass(store, "K1+", f1)).set_children(
            ref.GenParticle("K*0", recording_mass(store, "K*0", f2)).set_children(
                ref.GenParticle("K+", KAON_MASS), ref.GenParticle("pi-", PION_MASS)),
            ref.GenParticle("pi+", PION_MASS)),
        ref.GenParticle("gamma", 0.0))
    run("k1_gamma_callable", k1, n, 6, masses_store=store)

    # two decaying children side by side: 3 -> (2, 3) + a stable particle
    store = {}
    chain = ref.GenParticle("B", B0_MASS).set_children(
        ref.GenParticle("X", 1500.0).set_children(ref.GenParticle("a", PION_MASS), ref.GenParticle("b", KAON_MASS)),
        ref.GenParticle("Y", recording_mass(store, "Y", f1)).set_children(
            ref.GenParticle("c", PION_MASS), ref.GenParticle("d", PION_MASS), ref.GenParticle("e", 0.0)),
        ref.GenParticle("f", 105.6583755))
    run("two_branches", chain, n, (3 * 3 - 4) + 2 + (3 * 3 - 4), masses_store=store)

    # ---- second batch (appended: the draws of the cases above are unchanged) ----
    # the kernel switches from the step-by-step formulation to one accumulated Lorentz matrix at nine children
    for nb in (8, 9):
        run(f"flat{nb}", ref.nbody_decay(B0_MASS, [PION_MASS] * nb), n, n_draws_flat(nb))
    run("flat2_boost", ref.nbody_decay(B0_MASS, [PION_MASS, KAON_MASS]), n, 0 + 2, boost=boost_rows(n, 10))
    run("flat3_massless", ref.nbody_decay(100.0, [0.0, 0.0, 0.0]), n, n_draws_flat(3))
    run("kstar_gamma_fixed_boost_one_row", kstar_gamma(KSTARZ_MASS), n, 4, boost=boost_rows(1, 11))

    # tests/helpers/decays.py:47-76 with every mass fixed: recurse_w_max over two levels of fixed-mass intermediates
    k1_fixed = ref.GenParticle("B+", B0_MASS).set_children(
        ref.GenParticle("K1+", K1_MASS).set_children(
            ref.GenParticle("K*0", KSTARZ_MASS).set_children(
                ref.GenParticle("K+", KAON_MASS), ref.GenParticle("pi-", PION_MASS)),
            ref.GenParticle("pi+", PION_MASS)),
        ref.GenParticle("gamma", 0.0))
    run("k1_gamma_fixed", k1_fixed, n, 6)
    run("k1_gamma_fixed_boost", k1_fixed_copy(ref), n, 6, boost=boost_rows(n, 12))

    # a nine-body decay below the top particle (its products are boosted by the parent afterwards), with and without
    # a per-event boost of the whole event
    def nine_below(store_, frac_):
        return ref.GenParticle("A", 12000.0).set_children(
            ref.GenParticle("X", recording_mass(store_, "X", frac_)).set_children(
                *[ref.GenParticle(f"x{i}", PION_MASS) for i in range(9)]),
            ref.GenParticle("g", 0.0), ref.GenParticle("h", 938.272))

    f3 = np.clip(rng.beta(4.0, 4.0, n), 1e-6, 1.0)
    store = {}
    run("nine_body_subdecay", nine_below(store, f3), n, n_draws_flat(3) + n_draws_flat(9), masses_store=store)
    store = {}
    p = np.random.default_rng(13).normal(0.0, 30000.0, (n, 3))
    boost_a = np.concatenate([p, np.sqrt((p * p).sum(1, keepdims=True) + 12000.0**2)], axis=1)
    run("nine_body_subdecay_boost", nine_below(store, f3), n, n_draws_flat(3) + n_draws_flat(9), boost=boost_a,
        masses_store=store)

    # three levels, three-body decays inside: top -> R1(-> R2(-> a b c) d e) f, R1 callable, R2 fixed
    store = {}
    deep = ref.GenParticle("T", 9000.0).set_children(
        ref.GenParticle("R1", recording_mass(store, "R1", f2)).set_children(
            ref.GenParticle("R2", 2000.0).set_children(
                ref.GenParticle("a", PION_MASS), ref.GenParticle("b", KAON_MASS), ref.GenParticle("c", 0.0)),
            ref.GenParticle("d", PION_MASS), ref.GenParticle("e", 105.6583755)),
        ref.GenParticle("f", 938.272))
    run("three_levels_three_body", deep, n, 2 + n_draws_flat(3) + n_draws_flat(3), masses_store=store)
    return out


def k1_fixed_copy(ref):
    """A second instance of the fixed-mass K1 chain (a GenParticle generates with one boost_to setting per trace)."""
    return ref.GenParticle("B+", B0_MASS).set_children(
        ref.GenParticle("K1+", K1_MASS).set_children(
            ref.GenParticle("K*0", KSTARZ_MASS).set_children(
                ref.GenParticle("K+", KAON_MASS), ref.GenParticle("pi-", PION_MASS)),
            ref.GenParticle("pi+", PION_MASS)),
        ref.GenParticle("gamma", 0.0))


def main():
    install_tensorflow_stand_in()
    sys.path.insert(0, REFERENCE_SRC)
    import phasespace as ref  # the reference, unmodified

    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REFERENCE_SRC)), ref.__file__
    out = build(ref)
    np.savez_compressed(OUT, **out)
    print("written", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
