"""CPU oracle for the mask-driven MVDR hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in float64 numpy/scipy, the arithmetic the reference
repository performs inline in its scripts.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``real-time-audio-visual-zooming_b200/`` (the product) imports it, and the
product has no CPU fallback.

Parity pinning: the reference ships no tests or golden outputs for this path
(SURVEY.md section 4 / 8-C), so the oracle is pinned *differentially*: every
function here is checked in ``tests/test_oracle_golden.py`` against vectors
produced by importing and running the reference's own code
(``oracle/make_golden.py``, run in the authoring container where
``/root/reference`` exists; outputs committed under ``tests/golden/``).
The arithmetic primitives themselves live in third-party code the reference
calls and does not pin (scipy.signal.stft/istft -> pocketfft,
numpy.linalg.solve -> LAPACK gesv); this image has scipy 1.18.1 / numpy 2.3.5.
The streaming recursion (``streaming_mvdr``) has no counterpart in the
reference at all: "parity unpinned" for that one function.
"""
from .mvdr_oracle import *  # noqa: F401,F403
