"""nptranscript_b200 — B200-native (sm_100a CUDA) implementation of npTranscript's read→transcript-cluster hot path,
behind the IdentityProfileHolder call seam.  See DESIGN.md and include/npt.h."""
from . import abi  # noqa: F401

__all__ = ["abi"]
