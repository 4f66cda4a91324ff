"""Solver classes of the reference (``solver/``) re-hosted on the CUDA library: same names, kwargs and result keys."""
