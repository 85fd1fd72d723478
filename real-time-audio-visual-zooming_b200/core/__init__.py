"""Mirrors of the reference's importable modules under rt_av_zoom/core (same names, argument order, array
layouts and error behaviour), with the arithmetic running in the CUDA kernels of libavzoom.so."""
