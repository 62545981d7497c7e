"""GPU: generated regexes (tests/regex_generator.py, after the reference's RegexGenerator.java) through the C ABI against the
CPU oracle - bit exact on (matched, start, end) in all three modes, with `from` offsets, and for iterated find()."""
import numpy as np
import pytest

import needle_b200 as nb
from tests.oracle_lib import Oracle
from tests.regex_generator import RegexGenerator, print_node

pytestmark = pytest.mark.gpu


def batch_for(g, node, rng, n=96):
    strings = []
    for _ in range(n):
        s = g.generate_string(node)[:200]
        a = "".join(chr(int(c)) for c in rng.integers(32, 127, size=int(rng.integers(0, 12))))
        b = "".join(chr(int(c)) for c in rng.integers(32, 127, size=int(rng.integers(0, 12))))
        strings.append([s, a + s + b, a + b, s + s, ""][int(rng.integers(0, 5))])
    return strings


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_generated_regexes_match_the_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    checked = 0
    for _ in range(60):
        g = RegexGenerator(rng, int(rng.integers(1, 10)))
        node = g.generate()
        regex = print_node(node)
        flags = [0, nb.LEFTMOST_LONGEST, nb.CASE_INSENSITIVE][int(rng.integers(0, 3))]
        blob = nb.compile_to_bytes(regex, flags)
        pat, ora = nb.Pattern(blob, device=0), Oracle(blob)
        data, offsets, cw = nb.pack_haystacks(batch_for(g, node, rng))
        n = len(offsets) - 1
        for mode in (0, 1, 2):
            got = pat.match_batch(mode, data, offsets, cw)
            exp = ora.match_batch(mode, data, offsets, cw)
            for name, a, b in zip(("matched", "start", "end"), got, exp):
                if b is not None:
                    assert np.array_equal(a, b), (regex, hex(flags), mode, name, int(np.nonzero(a != b)[0][0]))
        lens = np.diff(offsets).astype(np.int64)
        frm = np.minimum(rng.integers(0, 6, size=n), np.maximum(lens - 1, 0)).astype(np.int32)
        got = pat.match_batch(2, data, offsets, cw, frm)
        exp = ora.match_batch(2, data, offsets, cw, frm)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b), (regex, hex(flags), "from")
        for a, b in zip(pat.find_all_batch(data, offsets, cw), ora.find_all_batch(data, offsets, cw)):
            assert np.array_equal(a, b), (regex, hex(flags), "find_all")
        pat.close()
        checked += 1
    assert checked == 60
