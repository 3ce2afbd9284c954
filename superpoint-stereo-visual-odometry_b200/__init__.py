"""spvo-frontend-b200: B200-native (sm_100a) SuperPoint decode + descriptor matching behind the
reference's front-end interface.  Compute lives in csrc/ (hand-written CUDA, C ABI declared in
include/spvo_frontend.h); this package is the host-side mirror of the reference interface."""
from ._lib import (DMATCH_DTYPE, KEYPOINT_DTYPE, MATCH_KNN_RATIO, MATCH_NN, MATCH_NN_CROSSCHECK, MATCHER_AUTO,
                   MATCHER_EXACT_FP32, MATCHER_TENSOR, LIB_PATH)
from .frontend import (CURR_LEFT, CURR_LEFT_CURR_RIGHT, CURR_LEFT_PREV_LEFT, CURR_RIGHT, PREV_LEFT,
                       PREV_LEFT_PREV_RIGHT, PREV_RIGHT, Frontend, SpvoError, SuperPointFeatureFrontEnd)

__all__ = [n for n in dir() if not n.startswith("_")]
