"""The numpy models of the kernels' index algebra (tools/) stay true: they are how the lane layouts, the even/odd split of
the 1024-point path, the overlap-add exchange, the mixed-radix row transform and the chirp-z path were validated before
the CUDA was written (no GPU needed)."""
import importlib.util
import os

import pytest

TOOLS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(TOOLS, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["fft1024_model"])
def test_model_runs(name, capsys):
    _load(name).main()
    assert "ok" in capsys.readouterr().out
