"""Models of the reference (``dyn/``) re-hosted on the CUDA library: same class names and attributes."""
