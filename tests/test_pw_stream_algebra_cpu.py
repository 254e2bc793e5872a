"""CPU: index algebra of csrc/pw_stream.cu (K / N permutations of the mma.sync fragments, ldmatrix.trans addressing of
the backward-weight kernel) replayed lane by lane in numpy -- scripts/emulate_pw_stream.py -- for every instantiated
shape.  Guards the permutation formulas against edits; the CUDA code itself is covered by the GPU tests."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("emulate_pw_stream", os.path.join(ROOT, "scripts", "emulate_pw_stream.py"))
emu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(emu)


@pytest.mark.parametrize("K,NO", [(16, 48), (48, 16), (32, 16), (16, 32), (24, 72), (72, 24), (40, 120), (8, 8)])
def test_forward_and_dgrad_fragment_permutations(K, NO):
    assert emu.fwd(K, NO, np.random.default_rng(K * 100 + NO)) < 1e-12


@pytest.mark.parametrize("Cin,Cout", [(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)])
def test_wgrad_ldmatrix_addressing(Cin, Cout):
    assert emu.wgrad(Cin, Cout, np.random.default_rng(Cin * 100 + Cout)) < 1e-12


def test_output_permutation_is_a_bijection_with_contiguous_lane_pieces():
    for NT in range(1, 19):
        seen = set()
        for j in range(NT):
            for tq in range(4):
                for e in range(2):
                    seen.add(emu.phys(NT, j, tq, e))
        assert seen == set(range(8 * NT))
        for q in range((NT + 3) // 4):
            R = emu.rq(NT, q)
            for tq in range(4):
                piece = [emu.phys(NT, 4 * q + jj, tq, e) for jj in range(R) for e in range(2)]
                assert piece == list(range(piece[0], piece[0] + 2 * R))         # one contiguous store per lane
