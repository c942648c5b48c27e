"""Depth of the witness program the engine runs (scripts/witness_depth.py replays engine.cu: do_open's substitutions on the
CPU): the native SHA-256 op and the regex state seeding must keep collapsing the long dependency chains - a regression in
the records the front end emits (circuit.hpp: ShaBlock, RegexSeed) shows up here as a jump in levels / iterations, long
before it costs milliseconds on the GPU."""
import os
import sys

import zkemail_b200 as z

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import witness_depth


def test_regex_template_collapses_to_a_few_levels():
    for style in (0, 1):
        c = z.Circuit("BodyHashRegex", [128, style])
        plain = witness_depth.analyse(c, True, False)
        seeded = witness_depth.analyse(c, True, True)
        assert plain[0] > 200 and seeded[0] <= 16, (style, plain[0], seeded[0])        # ~2-5 levels per byte -> a handful in total
        assert seeded[2] == plain[2]                                                  # every op is kept
        assert seeded[1] < plain[1]


def test_email_verifier_test_circuit_depth():
    c = z.Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1])
    generic = witness_depth.analyse(c, False, False)
    sha = witness_depth.analyse(c, True, False)
    both = witness_depth.analyse(c, True, True)
    assert generic[0] > sha[0] > both[0]
    assert both[0] <= 400                       # what is left is the Poseidon round chain of the public-key hash
    assert both[1] <= 1.25 * (-(-both[2] // witness_depth.T))    # within 25 % of the 512-ops-per-iteration floor
