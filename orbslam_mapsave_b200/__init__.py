"""orbslam_mapsave_b200 — B200-native (sm_100a) ORB front-end: ORBextractor + ORBmatcher Hamming searches.

The product is the CUDA library liborb_b200.so (csrc/, C ABI in include/orb_b200.h) and the C++ drop-in classes in
host/.  This Python package only binds the C ABI for tests and bench.py.
"""
from .capi import OrbError, KP_DTYPE, LIB_PATH, lib, device_count   # noqa: F401
from .extractor import ORBextractor                                  # noqa: F401
from . import matcher                                                                    # noqa: F401
from .matcher import ORBmatcher, FeatureVector, View, GridView, popc_peak, distinctive_descriptors      # noqa: F401
from .vocabulary import ORBVocabulary                                  # noqa: F401
from .map_archive import MapArchive                                   # noqa: F401
