"""oracle/_ref (C emitted by the reference's own code generator, when present) agrees with the
oracle port — guards the argument marshalling of the `kind: "reference"` CPU baseline."""
import os

import pytest

from oracle import oracle as O
from oracle import refrun
from helpers import iso_problem, rel_linf


@pytest.mark.parametrize('so', [8, 12])
def test_reference_generated_code_matches_oracle(so):
    ref = refrun.load_forward(so)
    if ref is None:
        pytest.skip("oracle/_ref not generated in this checkout")
    p = iso_problem(24, 8, so, 100.0)
    q = iso_problem(24, 8, so, 100.0)
    O.iso_forward(p['u'], so, p['w'], p['dt'], 1, p['nt'] - 2, damp=p['damp'], vp=1.5,
                  src=p['src'], rec=p['rec'])
    refrun.run_forward(ref, q['u'], q['damp'], 1.5, q['dt'], 1, q['nt'] - 2, q['src'], q['rec'], so,
                       threads=min(4, os.cpu_count() or 1))
    assert rel_linf(q['u'], p['u']) < 1e-5
    assert rel_linf(q['rec']['data'], p['rec']['data']) < 1e-5
