"""promonet_b200: B200-native (sm_100a) hot path of ProMoNet

Drop-in entry points (same names and arguments as the reference package):
`promonet_b200.synthesize.from_features`, `promonet_b200.preprocess.from_audio`
and `promonet_b200.train`.  Everything numerical runs in libpromonet_b200.so
(hand-written CUDA behind the C ABI of include/promonet_b200.h); there is no
CPU fallback.
"""
from .config import *
from . import _lib
from . import edit
from . import evaluate
from . import load
from . import model
from . import preprocess
from . import synthesize
