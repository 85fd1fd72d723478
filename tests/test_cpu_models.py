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


def test_cluster_mixer_model(capsys):
    """The cluster-resident mixer as shipped (decimated sequences per CTA, fused exchange step, natural-order radix-25
    block, quarter-wise I/O) against numpy's FFT and the float64 oracle mixer (the model asserts), and the sector
    efficiency of its exchange step that the natural-order block buys."""
    _load("mixer_cluster_model").main()
    out = capsys.readouterr().out
    assert "mixer err" in out
    ratio = float(out.strip().splitlines()[-1].split(":")[-1])
    assert ratio < 1.5
