"""GPU parity on 40 seeded random trees with odd NEWICK syntax (polytomies, missing / zero /
negative / scientific lengths, quoting, comments, support labels; 3 to 119 nodes), against
vectors the unmodified reference produced for them (tests/golden/make_golden_fuzz.py) and the
oracle.  Same bars as test_gpu_parity.py: structure and MRCA ids bit-exact, quartet rows
bit-exact, distances within 1e-12 of the fp64 path sum (against the path's L1 norm) and within
the reference's fp32 accumulation of its raw output.

(File name sorts last on purpose: written after this round's GPU budget was spent, so it runs
after every test that has already been seen green on a B200.)"""
import os

import numpy as np
import pytest
from conftest import GOLDEN

import oracle as O
from suchtree_b200 import SuchTree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("geom", [(0, 0), (2, 1)])
def test_random_trees_against_reference_and_oracle(geom):
    z = np.load(os.path.join(GOLDEN, "fuzz_trees.npz"))
    for k in range(40):
        text = str(z["t%d__newick" % k])
        T = SuchTree(text, _block_shift=geom[0], _micro_shift=geom[1])
        ot = O.OracleTree.from_newick(text)
        assert np.array_equal(T._ft.parent, z["t%d__parent" % k]), text
        assert T.depth == int(z["t%d__depth" % k]), text
        pairs = z["t%d__pairs" % k]
        assert np.array_equal(T.common_ancestors_bulk(pairs[:50]), z["t%d__mrca" % k]), text
        got = T.distances_bulk(pairs)
        d64, l1 = ot.distances_f64(pairs, with_l1=True)
        assert np.all(np.abs(got - d64) <= 1e-12 * l1 + 1e-300), text
        assert np.all(np.abs(got - z["t%d__distance" % k]) <= 2e-7 * T.depth * l1 + 1e-300), text
        if "t%d__q" % k in z.files:
            assert np.array_equal(T.quartet_topologies_bulk(z["t%d__q" % k]), z["t%d__t" % k]), text
