"""The kernel variants selected by environment switches (read once per process, so each runs in a
subprocess) must all reproduce the oracle on the B200: warp-scope x passes everywhere / nowhere,
short z chunks of the product pass, spectral reuse off."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

VARIANTS = [
    ({"LESGO_XW": "0"}, "16,16,6", "deriv,convec,steps"),
    ({"LESGO_XW": "2"}, "64,48,6", "deriv,convec,press,steps"),
    ({"LESGO_XW": "2"}, "512,64,3", "deriv,convec"),
    ({"LESGO_XW": "1"}, "512,64,3", "deriv,convec,press"),
    ({"LESGO_XW": "3"}, "1024,32,3", "deriv,convec,press,steps"),
    ({"LESGO_PROD_CHUNK": "3"}, "128,64,8", "convec,steps"),
    ({"LESGO_REUSE": "0"}, "32,32,6", "steps,full"),
    ({"LESGO_XW2": "7"}, "512,64,3", "deriv,convec,press,steps,full"),    # two-stage x inverse on every x-inverse launch
    ({"LESGO_XW2": "0"}, "512,64,3", "convec,steps"),                     # ... and on none
    ({"LESGO_FILTER_PRUNE": "0"}, "64,64,8", "lasd,full"),                # test filters without the zero-column pruning
]


@pytest.mark.gpu
@pytest.mark.parametrize("env,grid,what", VARIANTS)
def test_variant_on_gpu(env, grid, what):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_check.py"), "--grid", grid, "--what", what],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "variant_check ok" in r.stdout, (env, r.stdout[-2000:], r.stderr[-2000:])
