"""image_matching_b200 -- B200-native (sm_100a) SuperPoint + SuperGlue inference path.

Drop-in for the reference's ``superglue.models.matching_test.Matching`` (and the two model
classes it owns).  See DESIGN.md / INTEGRATION.md at the repo root.
"""
from .matching import (Matching, MatchingOfficial, SuperPoint, SuperPointOfficial, SuperGlue,  # noqa: F401
                       knn_ratio_match)
from .registration import estimate_affine_partial_2d, warp_affine, register_pairs, resize_u8  # noqa: F401
from . import synth, lib  # noqa: F401

__all__ = ["Matching", "MatchingOfficial", "SuperPoint", "SuperPointOfficial", "SuperGlue", "knn_ratio_match", "estimate_affine_partial_2d", "warp_affine",
           "register_pairs", "resize_u8", "synth", "lib"]
