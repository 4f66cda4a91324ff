"""robust-nonlinear-mpc_b200 -- B200-native SCP / fast-SLS / nominal-QP inner solve of robust nonlinear MPC.

The directory name carries a hyphen (it mirrors the reference repository's name); import it through the alias
package ``rnmpc_b200`` at the repository root, or put this directory on ``sys.path`` to get drop-in
``dyn.*`` / ``solver.*`` modules with the reference's names (``from solver.SCP_SLS_jit import SCP_SLS``).

Everything numerical happens in ``librnmpc_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/rnmpc_b200.h``).  There is no CPU fallback: importing ``engine`` without the built library, or creating a
solver without a CUDA device, raises.
"""
from ._lib import LibraryNotBuilt, load_library  # noqa: F401

__all__ = ["LibraryNotBuilt", "load_library"]
