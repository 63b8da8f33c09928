#!/usr/bin/env python
"""Regenerates tests/golden/*.json.

Two kinds of vectors:
  reference.json  constants copied out of the reference's own tests / docs (file:line given per
                  entry) and known-answer vectors of the published algorithms of its external
                  crates (ChaCha20 keystream, tss 0.2 packed sharing) -- these PIN the oracle;
  oracle.json     outputs of the pinned oracle (oracle/sda_oracle.c) on seeded inputs, frozen so
                  that the GPU parity tests also run against committed bytes, not only against a
                  live oracle build.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
import util  # noqa: E402
from sda_b200 import LinearMaskingScheme as LMS  # noqa: E402
from sda_b200 import LinearSecretSharingScheme as LSS  # noqa: E402
from sda_b200 import params  # noqa: E402

reference = {
    "full_loop": {
        "source": "integration-tests/tests/full_loop.rs:11-67,113,148",
        "modulus": 433, "dimension": 4, "participants": 2, "input": [1, 2, 3, 4], "expected_positive": [2, 4, 6, 8],
        "sharing": {
            "additive": {"share_count": 3, "modulus": 433},
            "packed_shamir": {"secret_count": 3, "share_count": 8, "privacy_threshold": 4, "prime_modulus": 433,
                              "omega_secrets": 354, "omega_shares": 150},
        },
        "masking": {"none": {}, "full": {"modulus": 433},
                    "chacha": {"modulus": 433, "dimension": 4, "seed_bitsize": 128}},
    },
    "cli_walkthrough": {
        "source": "README.md:86,105-107,157; docs/simple-cli-example.sh:38-44,56",
        "sharing": {"share_count": 3, "modulus": 433}, "dimension": 10,
        "inputs": [list(range(10)), [0] * 10, [0, 1] * 5],
        "expected": [0, 2, 2, 4, 4, 6, 6, 8, 8, 10],
    },
    "chacha20_kat": {
        "source": "rand 0.3 chacha.rs test_rng_true_values == RFC 7539 block function, key 0^256, counter 0/1, nonce 0",
        "block0": "ade0b876 903df1a0 e56a5d40 28bd8653 b819d2bd 1aed8da0 ccef36a8 c70d778b "
                  "7c5941da 8d485751 3fe02477 374ad8b8 f4b8436a 1ca11815 69b687c3 8665eeb2".split(),
        "block1_head": "bee7079f 7a385155 7c97ba98 0d082d73".split(),
    },
    "tss_kat": {
        "source": "threshold-secret-sharing 0.2 packed.rs tests (PSS_4_8_3 / PSS_4_26_3), re-derived in SURVEY.md 8c",
        "prime": 433, "omega_secrets": 354, "secrets": [1, 2, 3], "randomness": [8, 8, 8, 8],
        "polynomial": [113, 51, 261, 267, 108, 432, 388, 112],
        "shares_omega150_n8": [91, 337, 88, 425, 336, 51, 395, 160],
        "shares_omega17_n26": [77, 230, 91, 286, 179, 337, 83, 212, 88, 406, 58, 425, 345, 350, 336, 430, 404, 51, 60,
                               305, 395, 84, 156, 160, 112, 422],
    },
    "gen_range_model": {
        "source": "rand 0.3 ChaChaRng::from_seed(&[1,2,3,4]) + gen_range(0, m) (SURVEY.md App. A.3)",
        "seed_words": [1, 2, 3, 4],
        "m433": [59, 358, 179, 210, 379, 368, 395, 356, 411, 270],
        "m2p61m1": [744479744108572534, 2063552701369210773, 1475773878734499814, 1684626750962375274],
    },
}


def L(a):
    return np.asarray(a).astype(object).tolist()


def sharing_cases():
    rng = np.random.default_rng(20260101)
    cases = []
    schemes = [
        ("additive3_433", LSS.Additive(3, 433), 10),
        ("additive1_433", LSS.Additive(1, 433), 5),
        ("additive5_p61", LSS.Additive(5, params.P61), 37),
        ("additive3_pgen", LSS.Additive(3, params.P61_GENERIC), 33),
        ("additive7_433", LSS.Additive(7, 433), 19),
        ("packed_ref_433", params.reference_test(), 4),
        ("packed_ref_433_long", params.reference_test(), 100),
        ("packed_cfg3", params.config3(), 50),
        ("packed_cfg4", params.config4(), 43),
        ("packed_cfg5", params.config5(), 29),
        ("packed_generic_shape", util.packed_scheme(params.P61, 2, 3, 6, O), 21),
        ("packed_generic_prime", util.packed_scheme(params.P61_GENERIC, 3, 2, 5, O), 31),
    ]
    for name, s, dim in schemes:
        m = s.c.modulus
        for kind in ("canonical", "signed"):
            if kind == "signed" and m < (1 << 40):
                secrets = rng.integers(-5 * m, 5 * m, size=dim, dtype=np.int64)
            else:
                secrets = util.rand_secrets(rng, dim, m, kind)
            for rounds in (20, 8):
                seed = util.seed_bytes(f"{name}/{kind}/{rounds}")
                shares = util.oracle_generate(O, s, secrets, seed, rounds)
                cases.append({"name": f"{name}/{kind}/r{rounds}", "scheme": [int(x) for x in (
                    s.c.kind, s.c.share_count, s.c.secret_count, s.c.privacy_threshold, s.c.modulus, s.c.omega_secrets,
                    s.c.omega_shares)], "rounds": rounds, "seed": seed.hex(), "secrets": L(secrets),
                    "shares_canonical": L(util.canon(O, m, shares))})
    return cases


def masking_cases():
    rng = np.random.default_rng(7)
    cases = []
    for name, ms, dim in [("full_433", LMS.Full(433), 13), ("full_p61", LMS.Full(params.P61), 21),
                          ("chacha_433", LMS.ChaCha(433, 4, 128), 4), ("chacha_p61", LMS.ChaCha(params.P61, 50, 128), 50),
                          ("chacha_pgen_256", LMS.ChaCha(params.P61_GENERIC, 17, 256), 17),
                          ("chacha_433_40bit", LMS.ChaCha(433, 9, 40), 9)]:
        secrets = util.rand_secrets(rng, dim, ms.c.modulus)
        seed = util.seed_bytes(name)
        mask, masked = O.mask(util.to_oracle_masking(O, ms), secrets, O.rng_from_seed_bytes(seed))
        cases.append({"name": name, "scheme": [int(x) for x in (ms.c.kind, ms.c.modulus, ms.c.dimension, ms.c.seed_bitsize)],
                      "seed": seed.hex(), "secrets": L(secrets), "mask": L(mask),
                      "masked_canonical": L(util.canon(O, ms.c.modulus, masked))})
    return cases


def main():
    with open(os.path.join(HERE, "reference.json"), "w") as f:
        json.dump(reference, f, indent=1)
    with open(os.path.join(HERE, "oracle.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/sda_oracle.c)", "sharing": sharing_cases(),
                   "masking": masking_cases()}, f)
    print("wrote reference.json, oracle.json")


if __name__ == "__main__":
    main()
