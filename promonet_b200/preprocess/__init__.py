from .core import *
from . import loudness
from . import penn
from . import spectrogram
from . import viterbi
