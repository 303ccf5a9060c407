"""DecayLanguage-format dicts of the reference's tests/fromdecay/example_decay_chains.py, written out by
hand because ``decaylanguage`` is not installed: they are what DecFileParser.build_decay_chains /
DecayChain.to_dict produce for tests/fromdecay/example_decays.dec."""

_PI0_MODES = [
    {"bf": 0.988228297, "fs": ["gamma", "gamma"], "model": "PHSP", "model_params": ""},
    {"bf": 0.011738247, "fs": ["e+", "e-", "gamma"], "model": "PI0_DALITZ", "model_params": ""},
    {"bf": 0.000033392, "fs": ["e+", "e+", "e-", "e-"], "model": "PHSP", "model_params": ""},
    {"bf": 0.000000065, "fs": ["e+", "e-"], "model": "PHSP", "model_params": ""},
]


def _pi0():
    return {"pi0": [dict(m, fs=list(m["fs"])) for m in _PI0_MODES]}


# D+ with a single decay mode; both decays carry zfit="relbw" (example_decay_chains.py:10-13)
dplus_single = {"D+": [{"bf": 1, "fs": ["K-", "pi+", "pi+", {"pi0": [{"bf": 1, "fs": ["gamma", "gamma"], "model": "",
                                                                       "model_params": "", "zfit": "relbw"}]}],
                        "model": "PHSP", "model_params": "", "zfit": "relbw"}]}

pi0_4branches = _pi0()

dplus_4grandbranches = {"D+": [{"bf": 1.0, "fs": ["K-", "pi+", "pi+", _pi0()], "model": "PHSP", "model_params": ""}]}

dstarplus_big_decay = {"D*+": [
    {"bf": 0.677, "fs": [{"D0": [{"bf": 1.0, "fs": ["K-", "pi+"], "model": "PHSP", "model_params": ""}]}, "pi+"],
     "model": "VSS", "model_params": ""},
    {"bf": 0.307, "fs": [{"D+": [{"bf": 1.0, "fs": ["K-", "pi+", "pi+", _pi0()], "model": "PHSP", "model_params": ""}]}, _pi0()],
     "model": "VSS", "model_params": ""},
    {"bf": 0.016, "fs": [{"D+": [{"bf": 1.0, "fs": ["K-", "pi+", "pi+", _pi0()], "model": "PHSP", "model_params": ""}]}, "gamma"],
     "model": "VSP_PWAVE", "model_params": ""},
]}
